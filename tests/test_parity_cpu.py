"""oracle/parity.py on the CPU: the comparison harness itself, fed with the reference's own outputs standing in for a
GPU result (must report zero mismatches), with perturbed outputs (must report them), and with an accurate double
precision solve of a trajectory on which the reference itself is unsound (must classify it as explained by the
reference's own error, against 40-digit arithmetic)."""
import types

import numpy as np

from cs_pathplan_b200 import workloads
from oracle import msnap_oracle as mo
from oracle import msnap_structured as st
from oracle import parity, ref


def _fake_result(wp, ns, B, cfg):
    rc = parity.to_ref_config(cfg)
    off = np.arange(B + 1, dtype=np.int64) * (ns + 1)
    counts, _, s = ref.generate_batch(off, wp, rc, nthreads=0, cap=400, kind="parity")
    t, co, md, it, vwf = ref.reweighted_solve_batch(off, wp, rc, nthreads=0, kind="parity")
    so = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rows = np.concatenate([s[b, :counts[b]] for b in range(B)])
    return types.SimpleNamespace(sample_offset=so, samples=rows, times=t, coeff=co, iters=it, max_dev=md, vw_final=vwf,
                                 best_s=None)


def test_harness_accepts_the_reference_and_sees_perturbations():
    B, ns = 12, 16
    wp, _ = workloads.cfg2(B=B)
    cfg = workloads.synthetic_config(4, "shipped")
    res = _fake_result(wp, ns, B, cfg)
    so = np.arange(B + 1, dtype=np.int64) * ns
    p = parity.batch_parity(res, wp, so, cfg)
    parity.assert_parity(p)
    assert p["checked"] == B == p["within_bars_of_reference"] and p["max_row_err_m"] == 0.0 and p["max_coeff_err"] == 0.0
    # a subset in a different order, coefficients on part of it
    p = parity.batch_parity(res, wp, so, cfg, picks=[7, 2, 11], n_coeff=2)
    parity.assert_parity(p)
    assert p["checked"] == 3 and p["coeff_checked"] == 2 and p["max_row_err_m"] == 0.0
    # perturbations are seen and NOT explained away: one row moved by 1e-3 m, one iteration count, one dropped row
    res.samples[res.sample_offset[5] + 3, 1] += 1e-3
    res.iters[1] += 1
    res.sample_offset = res.sample_offset.copy()
    res.sample_offset[-1] -= 1
    p = parity.batch_parity(res, wp, so, cfg)
    assert p["count_mismatch"] == 1 and p["unexplained"] == 2 and p["iters_mismatch"] == 1
    un = {u["trajectory"]: u for u in p["reference_unsound"]["unexplained"]}
    assert set(un) == {5, B - 1} and abs(un[5]["row_err_vs_reference_m"] - 1e-3) < 1e-12 and not un[B - 1]["same_count"]


def test_reference_unsound_trajectory_is_classified_against_exact_arithmetic():
    """cfg2 trajectory 128 has a 0.1 s segment among 4 s ones: the reference's dense inverse of M loses ~12 digits there
    (two builds of the unmodified reference differ from each other by 1e-3 relative), so an accurate solve is NOT within
    1e-8 of it -- but it is several hundred times closer to the exact solution than the reference is."""
    ns = 16
    wp_all, _ = workloads.cfg2()                      # the benchmark's batch (the generator depends on B)
    wp = wp_all[128 * (ns + 1): 130 * (ns + 1)]
    cfg = workloads.synthetic_config(4, "shipped")
    vel = acc = np.zeros((2, 3))
    t, co, rows, so, iters, vwf, md, bs = [], [], [], [0], [], [], [], []
    for b in range(2):
        p = wp[b * (ns + 1):(b + 1) * (ns + 1)]
        T = mo.allocate_time(p, cfg.V_avg, cfg.min_time_s)
        out = st.reweighted_structured(4, p, vel, acc, T, cfg.path_weight, cfg.vel_zero_weight)    # double precision
        c = np.array([[[float(v) for v in ax] for ax in seg] for seg in out["coeff"]])
        r, _, _, _ = mo.sample_polynomials(c.reshape(ns, -1), T, 4, cfg.sample_distance)
        t.append(T); co.append(c); rows.append(np.array(r)); so.append(so[-1] + len(r))
        iters.append(out["iters"]); vwf.append(out["vw_final"]); md.append(float(out["max_dev"])); bs.append(out["best_s"])
    res = types.SimpleNamespace(sample_offset=np.array(so), samples=np.concatenate(rows), times=np.concatenate(t),
                                coeff=np.concatenate(co), iters=np.array(iters, dtype=np.int32), max_dev=np.array(md),
                                vw_final=np.array(vwf), best_s=np.concatenate(bs))
    p = parity.batch_parity(res, wp, np.array([0, ns, 2 * ns]), cfg)
    parity.assert_parity(p)
    ex = p["reference_unsound"]
    assert ex["trajectories"] == 1 == ex["examined"] == p["explained_by_reference_error"] and p["within_bars_of_reference"] == 1
    assert ex["max_coeff_err_vs_exact"] <= 1e-6 and ex["max_reference_coeff_err_vs_exact"] > 1e-5
    assert ex["max_gpu_over_reference_coeff_err"] < 0.01
    assert p["worst_vs_reference"]["coeff_err"] > 1e-5 and p["worst_vs_reference"]["trajectory_coeff"] == 0
