#!/usr/bin/env python3
"""Generate tests/golden/bezier_golden.npz from the reference's own source (BUILD container only: needs /root/reference).

Every output comes from executing the UNMODIFIED /root/reference/math_util/bezier.cpp (math_util::Bezier::
GenerateTrajectoryMatrix, driven like UavPathPlanner::Bezier_3D, uavPathPlanning.cpp:4477-4505) and the patrol
post-processing helpers cut out of /root/reference/uavPathPlanning.cpp:118-206 (oracle/Makefile, oracle/bezier_wrapper.cpp).
The reference ships no recorded outputs for these, so these executions are the pin.

    make -C oracle && python tests/golden/make_bezier_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bezier_ref as br  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bezier_golden.npz")


def random_walk(rng, ns, sigma=10.0):
    p0 = rng.uniform(-100, 100, 3)
    return np.vstack([p0, p0 + np.cumsum(rng.normal(0, sigma, (ns, 3)), 0)])


def cases():
    rng = np.random.default_rng(20261019)
    out = [("uav31_0_d300", ref.UAV31_0_ENU, 300.0, 0.0), ("uav31_0_d300_minr", ref.UAV31_0_ENU, 300.0, 30.0),
           ("uav31_0_default_res", ref.UAV31_0_ENU[:, :] * 0.01, -1.0, 0.0)]
    for ns in (1, 2, 5, 16, 40):
        for sd, mr in ((1.0, 0.0), (0.37, 0.0), (2.5, 50.0)):
            out.append((f"rw_ns{ns}_sd{sd}_mr{mr}", random_walk(rng, ns), sd, mr))
    sharp = np.array([[0, 0, 10.0], [400, 0, 12], [400, 30, 15], [0, 30, 11], [0, 60, 10], [400, 60, 9]])   # U-turns: k search runs
    out.append(("uturns_minr", sharp, 5.0, 80.0))
    out.append(("uturns_free", sharp, 5.0, 0.0))
    dup = random_walk(rng, 6)
    dup[3] = dup[2] + [0.05, 0.0, 3.0]                   # d < 0.1: end-waypoint fallback in the middle (bezier.cpp:41, 175-179)
    out.append(("short_segment_mid", dup, 1.0, 0.0))
    dup0 = random_walk(rng, 4)
    dup0[1] = dup0[0] + [0.0, 0.01, 0.0]                 # ... and on the FIRST segment (its start point is then never emitted)
    out.append(("short_segment_first", dup0, 1.0, 0.0))
    out.append(("exact_division", np.array([[0, 0, 0.0], [3, 0, 0], [6, 0, 3]]), 0.25, 0.0))   # t lands on 1.0 within rounding
    out.append(("vertical_only", np.array([[5, 5, 0.0], [5, 5, 40], [5, 5.05, 80]]), 1.0, 0.0))
    return out


def patrol_cases():
    rng = np.random.default_rng(7)
    sq = np.array([[0, 0, 50.0], [300, 0, 52], [300, 200, 51], [0, 200, 49]])
    bow = np.array([[0, 0, 0.0], [100, 100, 0], [100, 0, 0], [0, 100, 0]])
    loops = {"square": sq, "bowtie": bow, "tiny": sq[:3] * 1e-7,
             "touching": np.array([[0, 0, 0.0], [10, 0, 0], [10, 10, 0], [5, 0, 0], [0, 10, 0]])}
    for i in range(12):                                   # wobbly closed loops, some self-intersecting
        n = int(rng.integers(5, 40))
        a = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = 100.0 + rng.normal(0, 45.0 if i % 2 else 8.0, n)
        pts = np.column_stack([r * np.cos(a), r * np.sin(a), rng.uniform(40, 60, n)])
        if i % 3 == 0:                                    # swap two vertices: the loop crosses itself
            j = int(rng.integers(0, n - 3))
            pts[[j, j + 2]] = pts[[j + 2, j]]
        loops[f"loop{i}"] = pts
    return loops


def main():
    blob, manifest = {}, []
    for name, path, sd, mr in cases():
        path = np.ascontiguousarray(path, dtype=np.float64)
        rows = br.generate(path, sd, mr)
        blob[f"{name}/path"], blob[f"{name}/rows"] = path, rows
        manifest.append(dict(name=name, sample_distance_override=sd, min_radius_arg=mr))
        print(f"{name:28s} n={path.shape[0]:3d} rows={rows.shape[0]:5d}")
    pman = []
    for name, poly in patrol_cases().items():
        poly = np.ascontiguousarray(poly, dtype=np.float64)
        closed = np.vstack([poly, poly[:1]])
        blob[f"patrol/{name}/polygon"] = poly
        for sp in (25.0, 7.5, 0.0) if poly.shape[0] < 6 else (25.0,):
            blob[f"patrol/{name}/boundary_{sp}"] = br.sample_closed_polygon_boundary(poly, sp)
        pman.append(dict(name=name, self_intersection_closed=br.has_self_intersection(closed, True),
                         self_intersection_open=br.has_self_intersection(poly, False)))
        print(f"patrol {name:12s} n={poly.shape[0]:3d} closed-intersects={pman[-1]['self_intersection_closed']}")
    blob["manifest"] = np.frombuffer(json.dumps(dict(bezier=manifest, patrol=pman)).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
