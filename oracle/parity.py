"""oracle/parity.py -- TEST INFRASTRUCTURE, not product code.

Whole-batch parity of a GPU result against Oracle A (``oracle/ref.py`` = the unmodified reference
``/root/reference/math_util/minimum_snap.cpp`` compiled against the oracle's Eigen shim): every trajectory of the batch
goes through the reference's own ``GenerateTrajectoryMatrix`` (ms.cpp:22-206) and the sampled rows are compared one by
one; a subset also goes through the reference's reweighting loop around ``SolveQPClosedForm`` (ms.cpp:76-90, 227-649)
for the discrete decisions and the coefficients.

Used by ``tests/test_gpu_full_parity.py`` and by the ``cpu_baseline`` leg of ``bench.py`` (outside every timed region).
Bars (BASELINE.json north_star): sample counts equal, every row within 1e-6 m, reweighting iterations equal, segment
times bit-exact, coefficients within 1e-8 in the position-scaled metric (tests/helpers.scaled_coeff_err).
"""
from __future__ import annotations

import os

import numpy as np

from . import ref


def _scaled_coeff_err_per_traj(c, c_ref, T, seg_offset):
    m = c.shape[2]
    pw = T[:, None, None] ** np.arange(m - 1, -1, -1)[None, None, :]
    den = np.max(np.abs(c_ref) * pw, axis=2, keepdims=True)
    den[den == 0] = 1.0
    e = np.max(np.abs(c - c_ref) * pw / den, axis=(1, 2))                  # per segment
    return np.maximum.reduceat(e, seg_offset[:-1])


def to_ref_config(cfg) -> ref.RefConfig:
    return ref.RefConfig(order=int(cfg.order), path_weight=float(cfg.path_weight),
                         vel_zero_weight=float(cfg.vel_zero_weight), V_avg=float(cfg.V_avg),
                         min_time_s=float(cfg.min_time_s), sample_distance=float(cfg.sample_distance),
                         start_vel=tuple(cfg.start_vel), end_vel=tuple(cfg.end_vel), start_acc=tuple(cfg.start_acc),
                         end_acc=tuple(cfg.end_acc))


def batch_parity(res, wp, seg_offset, cfg, picks=None, n_coeff=512, kind="parity", threads=0, row_cap=None,
                 sample_distance_override=-1.0, v_avg_override=-1.0):
    """Compare trajectories ``picks`` (default: all) of the GPU result ``res`` (cs_pathplan_b200.api.BatchResult, or
    any object with sample_offset / samples / times / coeff / iters / max_dev / vw_final host arrays laid out for the
    whole batch) with the reference run on the same waypoints.

    Returns a dict: checked, count_mismatch, max_row_err_m, coeff_checked, iters_mismatch, time_mismatch,
    max_coeff_err, max_dev_err, seconds, threads, kind.  Nothing is asserted here."""
    import time

    seg_offset = np.asarray(seg_offset, dtype=np.int64)
    B = seg_offset.shape[0] - 1
    picks = np.arange(B) if picks is None else np.asarray(picks, dtype=np.int64)
    rc = to_ref_config(cfg)
    threads = threads or (os.cpu_count() or 1)
    # the picked trajectories as one CSR batch of their own
    p0 = seg_offset[picks] + picks
    n_pts = (seg_offset[picks + 1] - seg_offset[picks]) + 1
    pt_off = np.concatenate([[0], np.cumsum(n_pts)]).astype(np.int64)
    idx = np.concatenate([np.arange(a, a + n) for a, n in zip(p0, n_pts)])
    sub_wp = np.ascontiguousarray(wp[idx])
    so = np.asarray(res.sample_offset, dtype=np.int64)
    g_counts = (so[picks + 1] - so[picks]).astype(np.int64)
    cap = int(row_cap or (g_counts.max() + 8))
    t0 = time.perf_counter()
    counts, used, samples = ref.generate_batch(pt_off, sub_wp, rc, sample_distance_override, v_avg_override,
                                               nthreads=threads, cap=cap, kind=kind)
    counts = counts.astype(np.int64)
    bad = counts != g_counts
    max_row = 0.0
    worst = -1
    ok_idx = np.nonzero(~bad)[0]
    for i in ok_idx:
        b = picks[i]
        d = np.abs(res.samples[so[b]:so[b + 1]] - samples[i, :counts[i]])
        e = float(d.max()) if d.size else 0.0
        if not np.isfinite(e):
            e = float("inf")
        if e > max_row:
            max_row, worst = e, int(b)
    out = dict(checked=int(picks.shape[0]), count_mismatch=int(bad.sum()), max_row_err_m=max_row,
               worst_row_trajectory=worst, mismatched=[int(picks[i]) for i in np.nonzero(bad)[0][:16]],
               rows_checked=int(counts[~bad].sum()), threads=int(used), kind=kind)
    # discrete decisions + coefficients on the first n_coeff picks
    nc = int(min(n_coeff, picks.shape[0]))
    if nc > 0:
        sub = picks[:nc]
        t, co, md, it, vwf = ref.reweighted_solve_batch(pt_off[:nc + 1], sub_wp[:pt_off[nc]], rc, v_avg_override,
                                                        nthreads=threads, kind=kind)
        segs = np.concatenate([np.arange(seg_offset[b], seg_offset[b + 1]) for b in sub])
        loc_off = np.concatenate([[0], np.cumsum(seg_offset[sub + 1] - seg_offset[sub])]).astype(np.int64)
        g_t, g_c = np.asarray(res.times)[segs], np.asarray(res.coeff)[segs]
        err = _scaled_coeff_err_per_traj(g_c, co, t, loc_off)
        out.update(coeff_checked=nc, iters_mismatch=int(np.sum(np.asarray(res.iters)[sub] != it)),
                   vw_final_mismatch=int(np.sum(np.asarray(res.vw_final)[sub] != vwf)),
                   time_mismatch=int(np.sum(g_t != t)), max_coeff_err=float(err.max()),
                   max_dev_err=float(np.max(np.abs(np.asarray(res.max_dev)[sub] - md))))
    out["seconds"] = time.perf_counter() - t0
    return out


def assert_parity(p, coeff_tol=1e-8, row_tol=1e-6, max_dev_tol=1e-8):
    assert p["count_mismatch"] == 0, p
    assert p["max_row_err_m"] <= row_tol, p
    if "coeff_checked" in p:
        assert p["iters_mismatch"] == 0 and p["vw_final_mismatch"] == 0 and p["time_mismatch"] == 0, p
        assert p["max_coeff_err"] <= coeff_tol, p
        assert p["max_dev_err"] <= max_dev_tol, p
