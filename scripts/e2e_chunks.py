"""Developer tool: end-to-end (host buffers) throughput of cfg2 for different host-chunk counts."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads
B, NS = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 16
tool = TrajectoryGeneratorTool(0)
cfg = workloads.synthetic_config(4, "shipped")
wps = [torch.from_numpy(workloads.cfg2(B=B, seed=1234 + r)[0]).pin_memory().numpy() for r in range(4)]
cap = tool.sample_bound(cfg, wps[0], ns=NS)
n_seg = B * NS
out = {k: torch.empty(shape, dtype=dt).pin_memory().numpy() for k, shape, dt in (
    ("times", (n_seg,), torch.float64), ("coeff", (n_seg, 3, 8), torch.float64), ("max_dev", (B,), torch.float64),
    ("iters", (B,), torch.int32), ("vw_final", (B,), torch.float64), ("sample_offset", (B + 1,), torch.int64),
    ("samples", (cap, 3), torch.float64))}
out["flags"] = torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
out["best_s"] = torch.zeros(n_seg, dtype=torch.int32).pin_memory().numpy()
for chunks, zc in ((1, False), (2, False), (3, False), (4, False), (1, True)):
    tool.set_host_chunks(chunks)
    tool.set_zero_copy(zc)
    for i in range(5): tool.generate_batch(cfg, wps[i % 4], ns=NS, capacity=cap, out=out, stats=False)
    t0 = time.perf_counter()
    n = 40
    for i in range(n): tool.generate_batch(cfg, wps[i % 4], ns=NS, capacity=cap, out=out, stats=False)
    dt = (time.perf_counter() - t0) / n
    print(f"B={B} chunks={chunks} zero_copy={zc}: {dt*1e6:.0f} us/step  {B/dt/1e6:.2f} M traj/s", flush=True)
