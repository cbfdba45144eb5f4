"""Per-kernel times of the B = 1 call on the reference's own mission (cfg1: uav31_0, 6 segments, 168 samples)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads

wp1, cfg1, sdo1, vo1 = workloads.cfg1()
dev = torch.device("cuda")
with TrajectoryGeneratorTool(0) as tool:
    ns1 = wp1.shape[0] - 1
    for rep in (1, 8, 64):
        wp = np.tile(wp1, (rep, 1))
        d_wp = torch.from_numpy(wp).to(dev)
        cap = tool.sample_bound(cfg1, wp, ns=ns1, v_avg_override=vo1)
        off = torch.zeros(rep + 1, dtype=torch.int64, device=dev)
        rows = torch.zeros((cap, 3), dtype=torch.float64, device=dev)
        for _ in range(3):
            tool.generate_batch_dev(cfg1, d_wp, off, rows, ns=ns1, sample_distance_override=sdo1, v_avg_override=vo1)
        tool.synchronize()
        tool.profile_begin()
        for _ in range(20):
            tool.generate_batch_dev(cfg1, d_wp, off, rows, ns=ns1, sample_distance_override=sdo1, v_avg_override=vo1)
        prof = tool.profile_end()
        print(json.dumps({"B": rep, "samples": int(off[-1]), "kernels_us": {k: round(v["total_ms"] / 20 * 1e3, 1) for k, v in prof.items()}}))
