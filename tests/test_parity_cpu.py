"""oracle/parity.py on the CPU: the comparison harness itself, fed with the reference's own outputs standing in for a
GPU result (must report zero mismatches) and with perturbed outputs (must report them)."""
import types

import numpy as np

from cs_pathplan_b200 import workloads
from oracle import parity, ref


def _fake_result(wp, ns, B, cfg):
    rc = parity.to_ref_config(cfg)
    off = np.arange(B + 1, dtype=np.int64) * (ns + 1)
    counts, _, s = ref.generate_batch(off, wp, rc, nthreads=0, cap=400, kind="parity")
    t, co, md, it, vwf = ref.reweighted_solve_batch(off, wp, rc, nthreads=0, kind="parity")
    so = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rows = np.concatenate([s[b, :counts[b]] for b in range(B)])
    return types.SimpleNamespace(sample_offset=so, samples=rows, times=t, coeff=co, iters=it, max_dev=md, vw_final=vwf)


def test_harness_accepts_the_reference_and_sees_perturbations():
    B, ns = 12, 16
    wp, _ = workloads.cfg2(B=B)
    cfg = workloads.synthetic_config(4, "shipped")
    res = _fake_result(wp, ns, B, cfg)
    so = np.arange(B + 1, dtype=np.int64) * ns
    p = parity.batch_parity(res, wp, so, cfg, n_coeff=6)
    parity.assert_parity(p)
    assert p["checked"] == B and p["max_row_err_m"] == 0.0 and p["max_coeff_err"] == 0.0 and p["coeff_checked"] == 6
    # a subset in a different order
    p = parity.batch_parity(res, wp, so, cfg, picks=[7, 2, 11], n_coeff=3)
    parity.assert_parity(p)
    assert p["checked"] == 3 and p["max_row_err_m"] == 0.0
    # perturbations are seen: one row moved by 1e-3 m, one iteration count, one dropped row
    res.samples[res.sample_offset[5] + 3, 1] += 1e-3
    res.iters[1] += 1
    res.sample_offset = res.sample_offset.copy()
    res.sample_offset[-1] -= 1
    p = parity.batch_parity(res, wp, so, cfg, n_coeff=6)
    assert p["count_mismatch"] == 1 and p["mismatched"] == [B - 1]
    assert abs(p["max_row_err_m"] - 1e-3) < 1e-12 and p["worst_row_trajectory"] == 5 and p["iters_mismatch"] == 1
