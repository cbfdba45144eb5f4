#!/usr/bin/env python3
"""Time the batched altitude optimisation (k_alt_part by default; k_alt_prep + k_alt_solve* + ... with --policy 0 / 1; device-resident rows, per-kernel CUDA events on the
launching stream through msnap_profile_begin/end) on a cfg2-sized sampler output.  One JSON line.  The CPU leg (a
banded-Cholesky stand-in for the reference's per-trajectory SimplicialLDLT loop) is `python bench.py --impl rows-cpu`.

    python scripts/alt_bench.py [--B 4096] [--iters 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cs_pathplan_b200 import TrajectoryGeneratorTool, shipped_altitude_params, workloads  # noqa: E402


def workload(B, seed=4):
    return workloads.sampled_rows(B, seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--policy", type=int, default=2, help="msnap_set_altitude_policy: 2 partitioned (default), 0 lane pairs, 1 one lane")
    a = ap.parse_args()
    rows, off, elev = workload(a.B)
    n = rows.shape[0]
    p = shipped_altitude_params()
    dev = torch.device("cuda")
    d_rows0 = torch.from_numpy(rows).to(dev)
    d_rows = d_rows0.clone()
    d_off = torch.from_numpy(off).to(dev)
    d_elev = torch.from_numpy(elev).to(dev)
    d_solves = torch.zeros(a.B, dtype=torch.int32, device=dev)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6553.3
    with TrajectoryGeneratorTool(0) as tool:
        stream = torch.cuda.Stream()
        tool.set_altitude_policy(a.policy)
        tool.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            for _ in range(3):
                d_rows.copy_(d_rows0)
                tool.altitude_optimize_batch_dev(p, d_off, d_rows, d_elev, solves=d_solves)
            stream.synchronize()
            tool.profile_begin()
            for _ in range(a.iters):
                d_rows.copy_(d_rows0)
                tool.altitude_optimize_batch_dev(p, d_off, d_rows, d_elev, solves=d_solves)
            prof = tool.profile_end()
        tool.set_stream(None)
    ms = {k: v["total_ms"] / a.iters for k, v in prof.items()}
    total = sum(ms.values())
    solves = d_solves.cpu().numpy()
    out = {"workload": f"{a.B} sampled trajectories, 150-259 rows each ({n} rows), shipped altitude parameters "
                       f"(config.yaml:1-8), analytic terrain", "kernels_ms": ms, "ms_per_batch": total,
           "trajectories_per_s": a.B / (total * 1e-3), "rows_per_s": n / (total * 1e-3),
           "mean_pass2_solves": float(solves.mean()), "max_pass2_solves": int(solves.max()),
           "algorithmic_bytes_per_row": 40, "achieved_GBps": 40.0 * n / (total * 1e-3) / 1e9,
           "hbm_frac": 40.0 * n / (total * 1e-3) / 1e9 / hbm_peak,
           "bound": "latency: one dependent chain of n rows per solve and trajectory, (1 + solves) chains per trajectory"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
