"""oracle/parity.py -- TEST INFRASTRUCTURE, not product code.

Whole-batch parity of a GPU result against Oracle A (``oracle/ref.py`` = the unmodified reference
``/root/reference/math_util/minimum_snap.cpp`` compiled against the oracle's Eigen shim): every picked trajectory goes
through the reference's own ``GenerateTrajectoryMatrix`` (ms.cpp:22-206) and the sampled rows are compared one by one,
and through the reference's reweighting loop around ``SolveQPClosedForm`` (ms.cpp:76-90, 227-649) for the discrete
decisions (segment times, iterations, final weight) and the coefficients.

Bars (BASELINE.json north_star): sample counts equal, every row within 1e-6 m, coefficients within 1e-8 in the
position-scaled metric, decisions equal.

**Where the reference is not sound.**  The reference inverts the dense mapping matrix ``M`` whose entries span
``T^(2o-1)`` (ms.cpp:247-266, 350, 511); a trajectory that mixes a very short segment with long ones (cfg2: T = 0.1 s next
to 6 s, ~4 % of the batch) makes that inverse lose up to 12 digits, and two compilations of the unmodified reference
(with / without FMA contraction) then disagree with EACH OTHER by up to 1e-3 relative.  For a trajectory whose GPU result
is not within the bars of the reference, the same problem (same decisions) is therefore solved in 40-digit arithmetic
(``oracle/msnap_structured.py``, Oracle B) and the GPU must be within the bars of THAT -- or, for the few trajectories
so ill-conditioned that no double-precision solve reaches 1e-8 (cond * eps > 1e-8), at least as close to the exact
solution as the reference is.  The discrete worst-deviation decisions are checked against exact arithmetic too.  Both
populations are counted and reported, with the reference's own error beside the GPU's.

Used by ``tests/test_gpu_full_parity.py`` and by the ``cpu_baseline`` leg of ``bench.py`` (outside every timed region).
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import ref


def _scaled_coeff_err_per_traj(c, c_ref, T, seg_offset):
    m = c.shape[2]
    pw = T[:, None, None] ** np.arange(m - 1, -1, -1)[None, None, :]
    den = np.max(np.abs(c_ref) * pw, axis=2, keepdims=True)
    den[den == 0] = 1.0
    e = np.max(np.abs(c - c_ref) * pw / den, axis=(1, 2))                  # per segment
    e = np.where(np.isfinite(e), e, np.inf)
    return np.maximum.reduceat(e, seg_offset[:-1])


def to_ref_config(cfg) -> ref.RefConfig:
    return ref.RefConfig(order=int(cfg.order), path_weight=float(cfg.path_weight),
                         vel_zero_weight=float(cfg.vel_zero_weight), V_avg=float(cfg.V_avg),
                         min_time_s=float(cfg.min_time_s), sample_distance=float(cfg.sample_distance),
                         start_vel=tuple(cfg.start_vel), end_vel=tuple(cfg.end_vel), start_acc=tuple(cfg.start_acc),
                         end_acc=tuple(cfg.end_acc))


def _exact_one(job):
    """40-digit solve of one trajectory.  The worst-deviation decisions (ms.cpp:408-439) are taken in exact arithmetic and
    compared with the GPU's: `tie` = they differ only between samples whose squared deviations agree to 1e-9 relative
    (SURVEY.md section 7 hard part 2), `mismatch` = they differ otherwise.  The coefficients / rows returned are those of
    the GPU's decisions.  Returns (coeff [ns,3,2o] rounded to double, rows, tie, mismatch)."""
    import mpmath

    from . import msnap_oracle as mo
    from . import msnap_structured as st

    order, path, vel, acc, T, pw, vw_final, best_s, sd = job
    mpmath.mp.dps = 40
    truth = st.solve_structured(order, path, vel, acc, T, pw, vw_final, ctx=mpmath.mp)
    best_s = list(truth["best_s"]) if best_s is None else [int(v) for v in best_s]
    tie = mismatch = False
    if list(truth["best_s"]) != best_s:
        for k, (a, b) in enumerate(zip(best_s, truth["best_s"])):
            if a != b:
                da, db = truth["dist2"][k][a], truth["dist2"][k][b]
                if abs(da - db) > mpmath.mpf(1e-9) * max(abs(da), abs(db)):
                    mismatch = True
        tie = not mismatch
        truth = st.solve_structured(order, path, vel, acc, T, pw, vw_final, ctx=mpmath.mp, best_s=best_s)
    tc = np.array([[[float(v) for v in ax] for ax in seg] for seg in truth["coeff"]])
    rows, _, _, _ = mo.sample_polynomials(tc.reshape(len(T), -1), T, order, sd)
    return tc, np.array(rows).reshape(-1, 3), tie, mismatch


def _exact_many(jobs, threads):
    """Run the 40-digit solves in `threads` fresh interpreter processes (python -m oracle.parity <jobs.pkl> <i> <n> <out>).
    Plain subprocesses on purpose: the caller is a process with a live CUDA context, OpenMP worker threads and torch loaded,
    where fork()-based pools can deadlock; the children import only numpy + mpmath."""
    if not jobs:
        return []
    n = max(1, min(threads, len(jobs) // 2))
    if n == 1:
        return [_exact_one(j) for j in jobs]
    import pickle
    import subprocess
    import sys
    import tempfile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        jf = os.path.join(tmp, "jobs.pkl")
        with open(jf, "wb") as f:
            pickle.dump(jobs, f)
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        procs = [subprocess.Popen([sys.executable, "-m", "oracle.parity", jf, str(i), str(n), os.path.join(tmp, f"out{i}.pkl")],
                                  cwd=root, env=env) for i in range(n)]
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("oracle.parity exact-solve worker failed")
        out = [None] * len(jobs)
        for i in range(n):
            with open(os.path.join(tmp, f"out{i}.pkl"), "rb") as f:
                for k, r in pickle.load(f):
                    out[k] = r
        return out


def batch_parity(res, wp, seg_offset, cfg, picks=None, n_coeff=None, kind="parity", threads=0, row_cap=None,
                 sample_distance_override=-1.0, v_avg_override=-1.0, coeff_tol=1e-8, row_tol=1e-6, max_exact=400):
    """Compare trajectories ``picks`` (default: all) of the GPU result ``res`` (cs_pathplan_b200.api.BatchResult, or
    any object with sample_offset / samples / times / coeff / iters / max_dev / vw_final / best_s host arrays laid out for
    the whole batch) with the reference run on the same waypoints; ``n_coeff`` = how many of the picks (the first ones)
    also get the decision / coefficient comparison (default: all).  Nothing is asserted here: see ``assert_parity``."""
    seg_offset = np.asarray(seg_offset, dtype=np.int64)
    B = seg_offset.shape[0] - 1
    picks = np.arange(B) if picks is None else np.asarray(picks, dtype=np.int64)
    rc = to_ref_config(cfg)
    sd = sample_distance_override if sample_distance_override > 0 else rc.sample_distance
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
    row_err = np.zeros(picks.shape[0])
    for i in np.nonzero(~bad)[0]:
        b = picks[i]
        d = np.abs(res.samples[so[b]:so[b + 1]] - samples[i, :counts[i]])
        e = float(d.max()) if d.size else 0.0
        row_err[i] = e if np.isfinite(e) else np.inf
    row_out = (row_err > row_tol) | bad
    out = dict(checked=int(picks.shape[0]), rows_checked=int(counts[~bad].sum()), threads=int(used), kind=kind,
               count_mismatch_vs_reference=int(bad.sum()))
    # discrete decisions + coefficients
    nc = picks.shape[0] if n_coeff is None else int(min(n_coeff, picks.shape[0]))
    coeff_err = np.zeros(picks.shape[0])
    if nc > 0:
        sub = picks[:nc]
        t, co, md, it, vwf = ref.reweighted_solve_batch(pt_off[:nc + 1], sub_wp[:pt_off[nc]], rc, v_avg_override,
                                                        nthreads=threads, kind=kind)
        segs = np.concatenate([np.arange(seg_offset[b], seg_offset[b + 1]) for b in sub])
        loc_off = np.concatenate([[0], np.cumsum(seg_offset[sub + 1] - seg_offset[sub])]).astype(np.int64)
        g_t, g_c = np.asarray(res.times)[segs], np.asarray(res.coeff)[segs]
        coeff_err[:nc] = _scaled_coeff_err_per_traj(g_c, co, t, loc_off)
        md_err = np.abs(np.asarray(res.max_dev)[sub] - md)
        out.update(coeff_checked=nc, iters_mismatch=int(np.sum(np.asarray(res.iters)[sub] != it)),
                   vw_final_mismatch=int(np.sum(np.asarray(res.vw_final)[sub] != vwf)),
                   time_mismatch=int(np.sum(g_t != t)))
    coeff_out = coeff_err > coeff_tol
    # ---- trajectories outside the bars against the reference: against exact arithmetic instead
    outl = np.nonzero(row_out | coeff_out)[0]
    explained = np.zeros(picks.shape[0], dtype=bool)
    ex = dict(trajectories=int(outl.shape[0]), examined=0, max_coeff_err_vs_exact=0.0, max_row_err_vs_exact_m=0.0,
              count_mismatch_vs_exact=0, max_reference_coeff_err_vs_exact=0.0, max_reference_row_err_vs_exact_m=0.0,
              decision_ties=0, decision_mismatch=0, max_gpu_over_reference_coeff_err=0.0, unexplained=[])
    if outl.shape[0]:
        todo = outl[:max_exact]
        vel = np.array([rc.start_vel, rc.end_vel], dtype=float)
        acc = np.array([rc.start_acc, rc.end_acc], dtype=float)
        jobs = []
        for i in todo:
            b = picks[i]
            sl = slice(int(seg_offset[b]), int(seg_offset[b + 1]))
            jobs.append((rc.order, sub_wp[pt_off[i]:pt_off[i + 1]], vel, acc, np.asarray(res.times)[sl], rc.path_weight,
                         float(np.asarray(res.vw_final)[b]),
                         None if getattr(res, "best_s", None) is None else np.asarray(res.best_s)[sl], sd))
        for i, (tc, trows, tie, mism) in zip(todo, _exact_many(jobs, threads)):
            b = picks[i]
            sl = slice(int(seg_offset[b]), int(seg_offset[b + 1]))
            T = np.asarray(res.times)[sl]
            one = np.array([0, T.shape[0]], dtype=np.int64)
            g_e = float(_scaled_coeff_err_per_traj(np.asarray(res.coeff)[sl], tc, T, one)[0])
            g_rows = res.samples[so[b]:so[b + 1]]
            same = g_rows.shape == trows.shape
            r_e = float(np.abs(g_rows - trows).max()) if same else np.inf
            ex["examined"] += 1
            ex["decision_ties"] += int(tie)
            ex["decision_mismatch"] += int(mism)
            ex["count_mismatch_vs_exact"] += 0 if same else 1
            ex["max_coeff_err_vs_exact"] = max(ex["max_coeff_err_vs_exact"], g_e)
            if same:
                ex["max_row_err_vs_exact_m"] = max(ex["max_row_err_vs_exact_m"], r_e)
            ref_e, ref_r = None, None
            if i < nc:
                ref_e = float(_scaled_coeff_err_per_traj(co[loc_off[i]:loc_off[i + 1]], tc, T, one)[0])
                ex["max_reference_coeff_err_vs_exact"] = max(ex["max_reference_coeff_err_vs_exact"], ref_e)
            if not bad[i] and same:
                ref_r = float(np.abs(samples[i, :counts[i]] - trows).max())
                ex["max_reference_row_err_vs_exact_m"] = max(ex["max_reference_row_err_vs_exact_m"], ref_r)
            # explained: against exact arithmetic the GPU meets the bars -- or, where the problem itself is too
            # ill-conditioned for any double-precision solve to get there (a 0.1 s segment among 5 s ones: cond * eps >
            # 1e-8), is at least as close to the exact solution as the reference is
            ok = same and not mism and g_e <= max(coeff_tol, ref_e or 0.0) and r_e <= max(row_tol, ref_r or 0.0)
            if ref_e:
                ex["max_gpu_over_reference_coeff_err"] = max(ex["max_gpu_over_reference_coeff_err"], g_e / ref_e)
            explained[i] = ok
            if not ok and len(ex["unexplained"]) < 16:
                ex["unexplained"].append(dict(trajectory=int(b), coeff_err_vs_reference=float(coeff_err[i]),
                                              coeff_err_vs_exact=g_e, reference_coeff_err_vs_exact=ref_e,
                                              row_err_vs_reference_m=float(row_err[i]), row_err_vs_exact_m=r_e,
                                              reference_row_err_vs_exact_m=ref_r, same_count=bool(same)))
    sound = ~(row_out | coeff_out)
    out.update(
        # trajectories within the bars of the reference itself
        within_bars_of_reference=int(sound.sum()),
        max_row_err_m=float(row_err[sound].max()) if sound.any() else 0.0,
        max_coeff_err=float(coeff_err[sound].max()) if sound.any() else 0.0,
        max_dev_err=float(md_err[sound[:nc]].max()) if nc > 0 and sound[:nc].any() else 0.0,
        # the others: checked against 40-digit arithmetic; "explained" = within the bars of exact arithmetic, or at least
        # as close to it as the reference is
        reference_unsound=ex, explained_by_reference_error=int(explained.sum()),
        count_mismatch=int(np.sum(bad & ~explained)),
        unexplained=int(np.sum((row_out | coeff_out) & ~explained)),
        worst_vs_reference=dict(row_err_m=float(row_err[~bad].max()) if (~bad).any() else 0.0,
                                coeff_err=float(coeff_err.max()),
                                trajectory_rows=int(picks[int(np.argmax(np.where(bad, -1.0, row_err)))]),
                                trajectory_coeff=int(picks[int(np.argmax(coeff_err))])),
        bars=dict(coeff=coeff_tol, row_m=row_tol))
    out["seconds"] = time.perf_counter() - t0
    return out


def assert_parity(p, max_dev_tol=1e-8):
    assert p["unexplained"] == 0 and p["count_mismatch"] == 0, p
    assert p["reference_unsound"]["examined"] == p["reference_unsound"]["trajectories"], p
    assert p["max_row_err_m"] <= p["bars"]["row_m"] and p["max_coeff_err"] <= p["bars"]["coeff"], p
    if "coeff_checked" in p:
        assert p["iters_mismatch"] == 0 and p["vw_final_mismatch"] == 0 and p["time_mismatch"] == 0, p
        assert p["max_dev_err"] <= max_dev_tol, p


if __name__ == "__main__":  # exact-solve worker: jobs i, i + n, i + 2n, ... of the pickled job list
    import pickle
    import sys

    jf, i, n, of = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    with open(jf, "rb") as f:
        all_jobs = pickle.load(f)
    res = [(k, _exact_one(all_jobs[k])) for k in range(i, len(all_jobs), n)]
    with open(of, "wb") as f:
        pickle.dump(res, f)
