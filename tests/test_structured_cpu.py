"""The "good CPU" baseline (oracle/structured_cpu.cpp = the library's own sequential kernel set compiled for the host) against
the golden vectors generated from the reference's source.  Besides pinning the baseline that bench.py times, this runs the
library's row assembly, block elimination, reweighting loop and sampler -- the very device functions of
cs_pathplan_b200/csrc/msnap_{device,generic}.cuh -- on the CPU, where no GPU is needed to see a regression."""
import numpy as np
import pytest

from helpers import COEFF_TOL, SAMPLE_TOL, load_golden, oracle_cfg, scaled_coeff_err
from oracle import structured_ref as sr

CASES = load_golden()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_structured_cpu_matches_reference_golden(case):
    cfg = oracle_cfg(case.cfg)
    r = sr.generate_batch(case.path, cfg, ns=case.ns, sample_distance_override=case.sdo, v_avg_override=case.vo, threads=1)
    assert r["flags"][0] == 0
    assert np.array_equal(r["times"], case.time)                      # bit-exact time allocation
    assert int(r["iters"][0]) == case.iters and float(r["vw_final"][0]) == case.vw_final
    noise = case.ref_noise()
    assert scaled_coeff_err(r["coeff"].reshape(case.coeff.shape), case.coeff, case.time) <= COEFF_TOL + 4 * noise
    assert r["samples"].shape == case.samples.shape                   # same accept decisions
    assert np.max(np.abs(r["samples"] - case.samples)) <= SAMPLE_TOL
    assert abs(float(r["max_dev"][0]) - case.max_dev) <= 1e-8 + 4 * noise


def test_structured_cpu_batch_ragged_and_threads():
    """A ragged batch through the CSR indexing equals its members solved one by one; the thread count changes nothing."""
    from cs_pathplan_b200 import workloads

    wp, so = workloads.cfg5(B=40, seed=3, ns_min=1, ns_max=20)
    cfg = workloads.synthetic_config(4, "shipped")
    one = sr.generate_batch(wp, cfg, seg_offset=so, threads=1)
    many = sr.generate_batch(wp, cfg, seg_offset=so, threads=4)
    for k in ("times", "coeff", "max_dev", "iters", "vw_final", "sample_offset", "samples", "flags"):
        assert np.array_equal(one[k], many[k]), k
    for b in (0, 7, 39):
        ns = int(so[b + 1] - so[b])
        p0 = int(so[b]) + b
        single = sr.generate_batch(wp[p0:p0 + ns + 1], cfg, ns=ns, threads=1)
        assert np.array_equal(single["samples"], one["samples"][one["sample_offset"][b]:one["sample_offset"][b + 1]])
        assert np.array_equal(single["coeff"], one["coeff"][so[b]:so[b + 1]])


def test_structured_cpu_dense_output_counts_every_candidate():
    """sample_distance 0 keeps every candidate (the count pass takes its evaluation-free shortcut)."""
    from cs_pathplan_b200 import workloads
    from oracle import msnap_oracle as mo

    wp, ns = workloads.cfg2(B=3, ns=6, seed=9)
    cfg = workloads.synthetic_config(4, "plain", 0.0)
    r = sr.generate_batch(wp, cfg, ns=ns, threads=1)
    for b in range(3):
        samples, _ = mo.generate_trajectory_matrix(wp[b * (ns + 1):(b + 1) * (ns + 1)], cfg)
        mine = r["samples"][r["sample_offset"][b]:r["sample_offset"][b + 1]]
        assert mine.shape == samples.shape and np.max(np.abs(mine - samples)) <= SAMPLE_TOL
