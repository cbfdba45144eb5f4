"""Short device-resident run of the bench workload for ncu captures (never a source of bench numbers)."""
import argparse, sys
sys.path.insert(0, ".")
import torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--weights", default="shipped")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--ns", type=int, default=16)
ap.add_argument("--order", type=int, default=4)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--config", default="", help="cfg3 | cfg5 (BASELINE.json workloads at full size); default: --batch x --ns random walks")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tool = TrajectoryGeneratorTool(0)
stream = torch.cuda.Stream(device=dev)
tool.set_stream(stream.cuda_stream)
cfg = workloads.synthetic_config(a.order, a.weights, 0.0 if a.config == "cfg5" else 1.0)
B, ns, m = a.batch, a.ns, 2 * a.order
seg_off = None
if a.config == "cfg3":
    wp_h, ns = workloads.cfg3()
    B = wp_h.shape[0] // (ns + 1)
elif a.config == "cfg5":
    wp_h, so_h = workloads.cfg5()
    B, ns = so_h.shape[0] - 1, None
    seg_off = torch.from_numpy(so_h).to(dev)
else:
    wp_h = workloads.random_walks(B, ns, 1234)
cap = tool.sample_bound(cfg, wp_h, ns=ns, seg_offset=None if seg_off is None else so_h)
wp = torch.from_numpy(wp_h).to(dev)
n_seg = B * ns if ns else int(so_h[-1])
f64 = dict(dtype=torch.float64, device=dev)
times, coeff = torch.empty(n_seg, **f64), torch.empty(n_seg * 3 * m, **f64)
max_dev, vw = torch.empty(B, **f64), torch.empty(B, **f64)
iters = torch.empty(B, dtype=torch.int32, device=dev)
off = torch.empty(B + 1, dtype=torch.int64, device=dev)
samples = torch.empty((cap, 3), **f64)
flags = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(a.iters):
    tool.generate_batch_dev(cfg, wp, off, samples, ns=ns, seg_offset=seg_off, times=times, coeff=coeff, max_dev=max_dev, iters=iters,
                            vw_final=vw, flags=flags)
torch.cuda.synchronize()
print("rows", int(off[-1]), "mean iters", float(iters.double().mean()), "launches", tool.launch_count)
