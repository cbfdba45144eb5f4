#!/usr/bin/env python3
"""Generate tests/golden/planner_golden.npz from the reference's own statements (BUILD container only: needs /root/reference).

Every output comes from executing function definitions cut out of /root/reference/uavPathPlanning.cpp where it lies
(oracle/cut_planner.sh + oracle/planner_wrapper.cpp -> oracle/_ref/libplanner_ref.so): the WGS84 <-> ENU transforms
(cpp:894-1108), optimizeSegmentAltitudeENU (cpp:1329-1364 over optimizeHeights / optimizeHeightsGlobalSmooth, cpp:1575-1827)
and generateFollowerTrajectories with its four formation generators (cpp:3931-4398).  Stand-ins used by that build: the
oracle's Eigen shims (SimplicialLDLT = banded LDL' in natural order) and a numbers-only json.

    make -C oracle && python tests/golden/make_planner_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from alt_helpers import sampled_paths, terrain_grid  # noqa: E402
from oracle import geo  # noqa: E402
from oracle import planner_ref as pr  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "planner_golden.npz")


def leaders():
    rng = np.random.default_rng(11)
    out = {"uav31_0": ref.generate(ref.UAV31_0_ENU, ref.shipped_config(), 300.0, 30.0)}          # the shipped case's leader rows
    t = np.linspace(0, 1, 60)
    out["s_curve"] = np.column_stack([4000 * t, 600 * np.sin(4 * t), 150 + 30 * t])
    out["five_rows"] = np.column_stack([np.arange(5) * 50.0, np.arange(5) ** 2 * 10.0, np.full(5, 80.0)])   # N <= 5: no smoothing
    out["six_rows"] = np.column_stack([np.arange(6) * 50.0, np.arange(6) ** 2 * 10.0, np.full(6, 80.0)])
    out["two_rows"] = np.array([[0.0, 0.0, 10.0], [30.0, 40.0, 12.0]])
    out["one_row"] = np.array([[5.0, 6.0, 7.0]])
    hd = np.cumsum(rng.normal(0, 0.25, 200))
    out["wiggly"] = np.column_stack([np.cumsum(25 * np.cos(hd)), np.cumsum(25 * np.sin(hd)), 100 + np.cumsum(rng.normal(0, 0.5, 200))])
    return out


def main():
    blob, man = {}, dict(followers=[], altitude=[])
    origin = geo.README_ORIGIN
    starts = np.column_stack([109.56 + 0.001 * np.arange(9), 40.867 + 0.0005 * np.arange(9), 10.0 + np.arange(9)])
    for name, rows in leaders().items():
        blob[f"leader/{name}"] = rows
    cases = [("uav31_0", 1, 3, {}), ("uav31_0", 2, 4, {}), ("uav31_0", 3, 9, dict(cfg_max_row=4)), ("uav31_0", 4, 7, {}),
             ("uav31_0", 7, 2, {}),                                                     # unknown model = V shape
             ("s_curve", 1, 5, dict(in_formation_distance=80.0)), ("s_curve", 3, 9, dict(in_max_row=3)),
             ("s_curve", 4, 9, dict(cfg_formation_distance=1.0, cfg_position_misalignment=3.0)),    # distance clamped from below
             ("five_rows", 2, 3, {}), ("six_rows", 2, 3, {}), ("two_rows", 1, 2, {}), ("one_row", 3, 2, {}),
             ("wiggly", 1, 6, {}), ("wiggly", 4, 6, dict(in_uav_R=40.0, cfg_max_row=0))]
    for i, (leader, model, F, kw) in enumerate(cases):
        out = pr.followers(blob[f"leader/{leader}"], origin, model, starts[:F], **kw)
        blob[f"followers/{i}"] = out
        man["followers"].append(dict(leader=leader, model=model, n_followers=F, params=kw))
        print(f"followers {i:2d} {leader:10s} model {model} F {F} -> {out.shape}")
    blob["starts"] = starts
    grid, res, ox, oy = terrain_grid()
    for i, (seed, params, use_grid) in enumerate([(21, (1.0, 1.0, 0.3, 2.0, 10.0), True), (22, (1.0, 0.0, 2.0, 2.0, 50.0), True),
                                                  (23, (0.0, 2.0, 0.5, 2.0, 30.0), True), (24, (1.0, 1.0, 0.3, 2.0, 10.0), False)]):
        rows, off = sampled_paths(16, seed=seed)
        out, z1, ok = pr.altitude_batch(rows, off, params, grid if use_grid else None, res, ox, oy)
        blob[f"alt/{i}/rows"], blob[f"alt/{i}/off"], blob[f"alt/{i}/out"], blob[f"alt/{i}/z1"] = rows, off, out, z1
        man["altitude"].append(dict(params=params, grid=use_grid, resolution=res, origin_x=ox, origin_y=oy, ok=[int(v) for v in ok]))
        print(f"altitude {i} rows {rows.shape[0]} ok {ok.min()}")
    blob["manifest"] = np.frombuffer(json.dumps(man).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
