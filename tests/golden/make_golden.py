#!/usr/bin/env python3
"""Generate tests/golden/msnap_golden.npz from the reference's own source.

Runs in the BUILD container only (needs /root/reference): every output stored here comes from executing the
UNMODIFIED /root/reference/math_util/minimum_snap.cpp (compiled against oracle/shim/Eigen/Dense by oracle/Makefile,
parity build) on the inputs below.  The reference ships no recorded outputs of its own, so
these are the known-answer vectors every oracle and the CUDA path are pinned to.  For each case the file also holds
`truth_coeff`: the same optimisation problem solved in 60-digit arithmetic (oracle/msnap_structured.py), replaying
the reference's discrete decisions, which gives the reference's OWN rounding error for that case.

    make -C oracle && python tests/golden/make_golden.py
"""
import json
import os
import sys

import mpmath
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import msnap_oracle as mo  # noqa: E402
from oracle import msnap_structured as st  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "msnap_golden.npz")


def random_walk(rng, ns, sigma=10.0):
    p0 = rng.uniform(-100, 100, 3)
    return np.vstack([p0, p0 + np.cumsum(rng.normal(0, sigma, (ns, 3)), 0)])


def boustrophedon(rng, ns, lane_len=200.0, spacing=20.0, step=25.0, jitter=0.5):
    """cfg4-style patrol lanes."""
    per_lane = int(lane_len / step)
    pts = []
    lane, i = 0, 0
    while len(pts) < ns + 1:
        x = i * step if lane % 2 == 0 else lane_len - i * step
        pts.append([x, lane * spacing, 50.0 + 0.5 * lane * (1 if lane % 2 == 0 else -1)])
        i += 1
        if i > per_lane:
            i, lane = 0, lane + 1
    p = np.array(pts[: ns + 1])
    return p + rng.normal(0, jitter, p.shape)


def cases():
    out = []
    shipped = dict(order=2, path_weight=1e-7, vel_zero_weight=0.01, V_avg=200.0, min_time_s=1.0, sample_distance=300.0)
    # config 1: README waypoints, shipped YAML, getPlan's overrides (distance 300 m; leader_speed 30 and YAML 200)
    out.append(("uav31_0_v30", ref.UAV31_0_ENU, shipped, 300.0, 30.0))
    out.append(("uav31_0_v200", ref.UAV31_0_ENU, shipped, 300.0, 200.0))
    out.append(("uav31_0_planar", np.c_[ref.UAV31_0_ENU[:, :2], np.zeros(7)], shipped, 300.0, 30.0))  # Minisnap_EN
    rng = np.random.default_rng(20261018)
    bc = dict(start_vel=(1.0, 0.5, -0.2), end_vel=(0.3, -0.1, 0.0), start_acc=(0.0, 0.1, 0.0), end_acc=(0.1, 0.2, 0.3))
    for order in (2, 3, 4):
        for ns in (1, 2, 3, 8, 16):
            for wname, pw, vw in (("plain", 0.0, 0.0), ("shipped", 1e-7, 0.01), ("strong", 0.5, 0.3)):
                cfg = dict(order=order, path_weight=pw, vel_zero_weight=vw, V_avg=5.0, min_time_s=0.1, sample_distance=1.0)
                if (ns + order) % 2 == 0:
                    cfg.update(bc)
                out.append((f"rw_o{order}_ns{ns}_{wname}", random_walk(rng, ns), cfg, -1.0, -1.0))
    for ns in (1, 4):
        cfg = dict(order=5, path_weight=1e-7, vel_zero_weight=0.01, V_avg=5.0, min_time_s=0.1, sample_distance=1.0)
        out.append((f"rw_o5_ns{ns}_shipped", random_walk(rng, ns), cfg, -1.0, -1.0))
    # edge cases the reference accepts
    p = random_walk(rng, 6)
    p[3] = p[2]  # duplicated consecutive waypoint -> T = min_time_s, deviation ratio forced to 0 (ms.cpp:613-615)
    out.append(("dup_waypoint_o3", p, dict(order=3, path_weight=1e-7, vel_zero_weight=0.01, V_avg=5.0, min_time_s=0.1,
                                            sample_distance=1.0), -1.0, -1.0))
    line = np.outer(np.arange(6), [12.0, -5.0, 2.0]) + [3.0, 4.0, 5.0]  # collinear, equally spaced
    out.append(("collinear_o4", line, dict(order=4, path_weight=0.0, vel_zero_weight=0.0, V_avg=5.0, min_time_s=0.1,
                                            sample_distance=1.0), -1.0, -1.0))
    out.append(("short_T_o4", random_walk(rng, 5, 0.2), dict(order=4, path_weight=1e-7, vel_zero_weight=0.01, V_avg=5.0,
                                                              min_time_s=0.1, sample_distance=0.05), -1.0, -1.0))
    out.append(("dense10hz_o4", random_walk(rng, 7), dict(order=4, path_weight=0.0, vel_zero_weight=0.0, V_avg=5.0,
                                                           min_time_s=0.1, sample_distance=0.0), -1.0, -1.0))
    out.append(("vavg_zero_o3", random_walk(rng, 4), dict(order=3, path_weight=0.0, vel_zero_weight=0.0, V_avg=0.0,
                                                           min_time_s=2.0, sample_distance=1.0), -1.0, -1.0))
    # longer chains (dense reference cost grows as ns^3)
    out.append(("patrol_o4_ns64_shipped", boustrophedon(rng, 64), dict(order=4, path_weight=1e-7, vel_zero_weight=0.01,
                                                                        V_avg=5.0, min_time_s=0.1, sample_distance=1.0), -1.0, -1.0))
    out.append(("patrol_o4_ns128_plain", boustrophedon(rng, 128), dict(order=4, path_weight=0.0, vel_zero_weight=0.0,
                                                                        V_avg=5.0, min_time_s=0.1, sample_distance=1.0), -1.0, -1.0))
    out.append(("rw_o2_ns200_shipped", random_walk(rng, 200, 300.0), dict(order=2, path_weight=1e-7, vel_zero_weight=0.01,
                                                                           V_avg=30.0, min_time_s=1.0, sample_distance=300.0), -1.0, -1.0))
    return out


def main():
    mpmath.mp.dps = 60
    blob, manifest = {}, []
    for name, path, cfgd, sdo, vo in cases():
        rc = ref.RefConfig(**cfgd)
        path = np.ascontiguousarray(path, dtype=np.float64)
        samples = ref.generate(path, rc, sdo, vo)
        rw = ref.reweighted_solve(path, rc, vo)
        # the reference's discrete decisions (arg-max sample per segment) from the line-by-line port
        vw_hist = rc.vel_zero_weight
        for _ in range(rw.iters):
            vw_hist = 0.01 if vw_hist < 1e-6 else vw_hist * 2.0
        assert vw_hist == rw.vw_final
        Vel = np.array([rc.start_vel, rc.end_vel], dtype=float)
        Acc = np.array([rc.start_acc, rc.end_acc], dtype=float)
        _, sinfo = mo.solve_qp_closed_form(rc.order, path, Vel, Acc, rw.time, rc.path_weight, rw.vw_final)
        best_s = sinfo.best_s if rc.path_weight > 0 else np.zeros(len(rw.time), dtype=np.int64)
        truth = st.solve_structured(rc.order, path, Vel, Acc, rw.time, rc.path_weight, rw.vw_final, ctx=mpmath.mp,
                                    best_s=[int(v) for v in best_s])
        truth_coeff = np.array([[[float(v) for v in ax] for ax in seg] for seg in truth["coeff"]])
        blob[f"{name}/path"] = path
        blob[f"{name}/time"] = rw.time
        blob[f"{name}/coeff"] = rw.coeff
        blob[f"{name}/samples"] = samples
        blob[f"{name}/best_s"] = np.asarray(best_s, dtype=np.int64)
        blob[f"{name}/truth_coeff"] = truth_coeff
        manifest.append(dict(name=name, cfg=cfgd, sample_distance_override=sdo, v_avg_override=vo,
                             max_dev=rw.max_dev, iters=rw.iters, vw_final=rw.vw_final,
                             truth_max_dev=float(truth["max_dev"])))
        print(f"{name:28s} ns={len(rw.time):4d} S={samples.shape[0]:5d} iters={rw.iters:2d} max_dev={rw.max_dev:.6f}")
    blob["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
