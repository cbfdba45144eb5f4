"""Box ceiling of the end-to-end curve: N ranks (one per GPU) each copy `--mb` MB device -> pinned host memory back to
back, all at once; reports GB/s per rank and in total.  The e2e figure of bench.py moves ~20 MB (samples only) or ~33 MB
(all outputs) device -> host per 4 096-trajectory step, so  ceiling_GBps / bytes_per_step  bounds its trajectories/s.

    python scripts/d2h_ceiling.py                                            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/d2h_ceiling.py
Options: --pin 1 pins the rank's host threads (and so its pinned pages, first touch) to the GPU's NUMA node first."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from cs_pathplan_b200.hostpin import pin_to_gpu_numa

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=33.0)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--pin", type=int, default=0)
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
pinned_to = pin_to_gpu_numa(local) if a.pin else None
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
n = int(a.mb * 1e6) // 8
d = torch.empty(n, dtype=torch.float64, device=dev).normal_()
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.zero_()                                   # first touch
up = torch.empty(int(1.7e6) // 8, dtype=torch.float64).pin_memory()
dup = torch.empty_like(up, device=dev)
out = {}
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)),
                 ("d2h_plus_h2d", lambda: (h.copy_(d, non_blocking=True), dup.copy_(up, non_blocking=True)))):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = n * 8 * a.iters / dt / 1e9
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per = [float(v.item()) for v in allv]
    else:
        per = [gbs]
    out[name] = {"per_rank_GBps": per, "total_GBps": sum(per), "min_GBps": min(per)}
if rank == 0:
    line = json.dumps({"n_gpus": world, "mb_per_copy": a.mb, "numa_pinned": bool(a.pin), "cpus": pinned_to and len(pinned_to), **out})
    print(line)
    if a.out:
        with open(a.out, "a") as f:
            f.write(line + "\n")
if world > 1:
    dist.destroy_process_group()
