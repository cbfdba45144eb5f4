"""Developer experiment: device-resident throughput of cfg2 when consecutive steps alternate between S handles/streams."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = 8
NSETS = 8 if B <= 65536 else 2
dev = torch.device("cuda", 0)
cfg = workloads.synthetic_config(4, "shipped")
for S in (1, 2):
    tools = [TrajectoryGeneratorTool(0) for _ in range(S)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    for t, s in zip(tools, streams): t.set_stream(s.cuda_stream)
    sets = []
    for r in range(NSETS):
        wp_h, _ = workloads.cfg2(B=B, ns=NS, seed=1234 + r)
        cap = tools[0].sample_bound(cfg, wp_h, ns=NS)
        f64 = dict(dtype=torch.float64, device=dev)
        sets.append(dict(wp=torch.from_numpy(wp_h).to(dev), off=torch.empty(B + 1, dtype=torch.int64, device=dev),
                         samples=torch.empty((cap, 3), **f64), times=torch.empty(B * NS, **f64),
                         coeff=torch.empty(B * NS * 3 * m, **f64), max_dev=torch.empty(B, **f64), vw=torch.empty(B, **f64),
                         iters=torch.empty(B, dtype=torch.int32, device=dev), flags=torch.empty(B, dtype=torch.int32, device=dev)))
    def step(i):
        s = sets[i % NSETS]
        tools[i % S].generate_batch_dev(cfg, s["wp"], s["off"], s["samples"], ns=NS, times=s["times"], coeff=s["coeff"],
                                        max_dev=s["max_dev"], iters=s["iters"], vw_final=s["vw"], flags=s["flags"])
    for i in range(24): step(i)
    torch.cuda.synchronize()
    K = 1200 if B <= 65536 else 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = [torch.cuda.Event() for _ in range(S)]
    e0.record(streams[0])
    for st in streams[1:]: st.wait_event(e0)
    for i in range(K): step(i)
    for k in range(1, S):
        done[k].record(streams[k]); streams[0].wait_event(done[k])
    e1.record(streams[0]); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"streams={S}: {ms / K * 1e3:.1f} us/step  {B * K / ms / 1e3:.2f} M traj/s", flush=True)
    for t in tools: t.close()
