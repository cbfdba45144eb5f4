"""Device-resident throughput of the BASELINE.json configurations other than the headline one (cfg3, cfg4, cfg5) and of
cfg2 for comparison.  Developer/measurement tool: prints one JSON line per configuration (CUDA-event timing on the
solver's stream, inputs resident in HBM, outputs to HBM)."""
import argparse, json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5")
ap.add_argument("--weights", default="shipped,plain")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (smoke runs)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tool = TrajectoryGeneratorTool(0)
stream = torch.cuda.Stream(device=dev)
tool.set_stream(stream.cuda_stream)


def make(name):
    if name == "cfg2":
        wp, ns = workloads.cfg2(B=int(4096 * a.scale)); return wp, ns, None, 1.0
    if name == "cfg3":
        wp, ns = workloads.cfg3(B=int((1 << 20) * a.scale)); return wp, ns, None, 1.0
    if name == "cfg4":
        wp, ns = workloads.cfg4(B=int(1024 * a.scale)); return wp, ns, None, 1.0
    wp, so = workloads.cfg5(B=int(65536 * a.scale)); return wp, None, so, 0.0


for name in a.configs.split(","):
    wp_h, ns, so_h, sd = make(name)
    for weights in a.weights.split(","):
        cfg = workloads.synthetic_config(4, weights, sample_distance=sd)
        B = (wp_h.shape[0] // (ns + 1)) if ns else so_h.shape[0] - 1
        n_seg = B * ns if ns else int(so_h[-1])
        cap = tool.sample_bound(cfg, wp_h, ns=ns, seg_offset=so_h)
        wp = torch.from_numpy(wp_h).to(dev)
        so = None if so_h is None else torch.from_numpy(so_h).to(dev)
        f64 = dict(dtype=torch.float64, device=dev)
        coeff = torch.empty(n_seg * 24, **f64); times = torch.empty(n_seg, **f64)
        max_dev, vw = torch.empty(B, **f64), torch.empty(B, **f64)
        iters = torch.empty(B, dtype=torch.int32, device=dev); flags = torch.empty(B, dtype=torch.int32, device=dev)
        off = torch.empty(B + 1, dtype=torch.int64, device=dev); samples = torch.empty((cap, 3), **f64)

        def step():
            tool.generate_batch_dev(cfg, wp, off, samples, ns=ns, seg_offset=so, times=times, coeff=coeff, max_dev=max_dev,
                                    iters=iters, vw_final=vw, flags=flags)
        for _ in range(2): step()
        torch.cuda.synchronize()
        l0 = tool.launch_count
        tool.profile_begin()
        step()
        prof = tool.profile_end()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps): step()
        e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        rows = int(off[-1].item())
        print(json.dumps({"config": name, "weights": weights, "B": B, "segments": n_seg, "ms_per_step": ms,
                          "trajectories_per_s": B / ms * 1e3, "segments_per_s": n_seg / ms * 1e3, "rows": rows,
                          "mean_iters": float(iters.double().mean().item()), "flagged": int((flags != 0).sum().item()),
                          "kernels_ms": {k: round(v["total_ms"], 4) for k, v in prof.items()}}), flush=True)
        del wp, coeff, times, samples, off
        torch.cuda.empty_cache()
