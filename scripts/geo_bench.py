#!/usr/bin/env python3
"""Time the batched WGS84 <-> ENU kernels (device-resident rows, CUDA events on the launching stream).  One JSON line per
kernel.  The CPU leg (oracle/geo_port.c on the host cores) is `python bench.py --impl rows-cpu`.

    python scripts/geo_bench.py [--rows 16777216] [--iters 20]
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
from cs_pathplan_b200 import TrajectoryGeneratorTool  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 24)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    n = a.rows
    ref = np.array([109.56059880227296, 40.86719901015758, 0.0])
    g = torch.Generator(device="cuda").manual_seed(11)
    enu = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    enu[:, :2] = torch.randn((n, 2), generator=g, dtype=torch.float64, device="cuda") * 2.0e4
    enu[:, 2] = torch.rand(n, generator=g, dtype=torch.float64, device="cuda") * 5000.0
    lla = torch.empty_like(enu)
    back = torch.empty_like(enu)
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6553.3
    with TrajectoryGeneratorTool(0) as tool:
        stream = torch.cuda.Stream()
        tool.set_stream(stream.cuda_stream)
        fp64_peak = tool.measure_fp64_peak()
        tool.enu_to_wgs84_dev(ref, enu, lla, steps_out=steps)
        stream.synchronize()
        mean_steps = float(steps.float().mean())
        out = []
        for name, fn in (("enu_to_wgs84", lambda: tool.enu_to_wgs84_dev(ref, enu, lla)),
                         ("wgs84_to_enu", lambda: tool.wgs84_to_enu_dev(ref, lla, back)),
                         ("enu_to_wgs84<exact_trig>", lambda: tool.enu_to_wgs84_dev(ref, enu, lla))):
            tool.set_geo_exact_trig(name.endswith("<exact_trig>"))
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(a.iters):
                    fn()
                e1.record(stream)
            stream.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            gbs = 48.0 * n / (ms * 1e-3) / 1e9
            out.append({"kernel": "k_" + name, "rows": n, "ms": ms, "rows_per_s": n / (ms * 1e-3),
                        "algorithmic_bytes_per_row": 48, "achieved_GBps": gbs, "hbm_peak_GBps": hbm_peak,
                        "hbm_frac": gbs / hbm_peak, "fp64_dfma_peak_TFLOPs": fp64_peak,
                        "mean_fixed_point_steps": mean_steps if name.startswith("enu_to_wgs84") else None,
                        "input": "rows 402 MB >> L2 (126 MB)"})
        tool.set_geo_exact_trig(False)
        tool.set_stream(None)
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
