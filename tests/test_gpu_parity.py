"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(cs_pathplan_b200/libmsnap_b200.so); the oracle is only the checker.

Bars (BASELINE.json north_star): coefficients within 1e-8 relative (position-scaled metric, helpers.scaled_coeff_err;
the reference's OWN rounding error against 60-digit arithmetic is added where it exceeds the bar, and the GPU result is
held to 1e-8 against that exact solution unconditionally), every sampled ENU position within 1e-6 m, identical
discrete decisions: segment times bit-exact, arg-max sample per segment, reweighting iterations, sample counts."""
import numpy as np
import pytest

from cs_pathplan_b200 import MinimumSnapConfig, shipped_config, workloads
from cs_pathplan_b200._lib import MsnapError
from helpers import COEFF_TOL, SAMPLE_TOL, load_golden, make_cfg, oracle_cfg, scaled_coeff_err

pytestmark = pytest.mark.gpu
CASES = load_golden()


def gpu_cfg(d):
    return make_cfg(MinimumSnapConfig, d)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_generate_matches_reference_golden(tool, case):
    res = tool.generate_batch(gpu_cfg(case.cfg), case.path, ns=case.ns, sample_distance_override=case.sdo,
                              v_avg_override=case.vo)
    assert res.flags[0] == 0
    assert np.array_equal(res.times, case.time)                                   # bit-exact time allocation
    assert res.iters[0] == case.iters and res.vw_final[0] == case.vw_final        # reweighting decisions
    if case.cfg["path_weight"] > 0 and case.ns > 1:
        assert np.array_equal(res.best_s, case.best_s)                            # arg-max decisions
    noise = case.ref_noise()
    assert scaled_coeff_err(res.coeff, case.truth_coeff, case.time) <= COEFF_TOL  # vs exact arithmetic
    assert scaled_coeff_err(res.coeff, case.coeff, case.time) <= COEFF_TOL + 4 * noise
    s = res.trajectory(0)
    assert s.shape == case.samples.shape                                           # sample count
    assert np.max(np.abs(s - case.samples)) <= SAMPLE_TOL
    assert abs(res.max_dev[0] - case.max_dev) <= 1e-8 + 4 * noise


def test_reference_method_signatures(tool):
    """The two reference methods, called the way Minisnap_3D / a SolveQPClosedForm user would."""
    c1 = next(c for c in CASES if c.name == "uav31_0_v30")
    s = tool.GenerateTrajectoryMatrix(c1.path, shipped_config(), 300.0, 30.0)
    assert s.shape == (168, 3) and np.max(np.abs(s - c1.samples)) <= SAMPLE_TOL
    c = next(c for c in CASES if c.name == "rw_o4_ns8_plain")
    Vel = np.array([c.cfg.get("start_vel", (0, 0, 0)), c.cfg.get("end_vel", (0, 0, 0))], dtype=float)
    Acc = np.array([c.cfg.get("start_acc", (0, 0, 0)), c.cfg.get("end_acc", (0, 0, 0))], dtype=float)
    poly, md = tool.SolveQPClosedForm(4, c.path, Vel, Acc, c.time, 0.0, 0.0, return_max_deviation=True)
    assert poly.shape == (8, 24) and md == 0.0
    assert scaled_coeff_err(poly.reshape(8, 3, 8), c.coeff, c.time) <= COEFF_TOL + 4 * c.ref_noise()


def test_error_behaviour_like_the_reference(tool):
    # fewer than 2 waypoints / fewer than 3 columns: empty matrix (ms.cpp:54-57)
    assert tool.GenerateTrajectoryMatrix(np.zeros((1, 3)), MinimumSnapConfig()).shape == (0, 0)
    assert tool.GenerateTrajectoryMatrix(np.zeros((5, 2)), MinimumSnapConfig()).shape == (0, 0)
    # orders the reference cannot represent (int overflow at >= 6) or that have no free derivative (1) are rejected
    for order in (0, 1, 6):
        with pytest.raises(MsnapError):
            tool.generate_batch(MinimumSnapConfig(order=order), np.random.rand(3, 3), ns=2)
    with pytest.raises(ValueError):
        tool.generate_batch(MinimumSnapConfig(), np.random.rand(7, 3), ns=3)      # 7 rows is not a multiple of 4


def test_capacity_overflow_is_reported_with_exact_layout(tool):
    wp, ns = workloads.cfg2(B=8, ns=4)
    cfg = workloads.synthetic_config(4, "plain")
    full = tool.generate_batch(cfg, wp, ns=ns)
    with pytest.raises(MsnapError) as e:
        tool.generate_batch(cfg, wp, ns=ns, capacity=int(full.sample_offset[3]) + 2)
    part = e.value.partial
    assert np.array_equal(part.sample_offset, full.sample_offset)                 # layout is still exact
    assert np.array_equal(part.samples, full.samples[: part.samples.shape[0]])    # rows that fit are right
    assert (part.flags[:3] & 2).max() == 0 and (part.flags[4:] & 2).min() == 2


@pytest.mark.parametrize("order", [2, 3, 4, 5])
@pytest.mark.parametrize("weights", ["plain", "shipped"])
def test_batch_vs_oracle_port_seeded(tool, order, weights):
    """A seeded mixed-length batch through the CSR entry point against the line-by-line port."""
    from oracle import msnap_oracle as mo

    rng = np.random.default_rng(100 * order + len(weights))
    lens = [1, 2, 3, 5, 8, 13, 4] if order < 5 else [1, 2, 3, 4]
    paths = []
    for n in lens:
        p0 = rng.uniform(-100, 100, 3)
        paths.append(np.vstack([p0, p0 + np.cumsum(rng.normal(0, 10, (n, 3)), 0)]))
    so = np.concatenate([[0], np.cumsum(lens)])
    cfg = workloads.synthetic_config(order, weights)
    cfg.start_vel, cfg.end_acc = (0.5, -0.25, 0.1), (0.0, 0.3, -0.1)
    res = tool.generate_batch(cfg, np.vstack(paths), seg_offset=so)
    ocfg = oracle_cfg({k: getattr(cfg, k) for k in ("order", "path_weight", "vel_zero_weight", "V_avg", "min_time_s",
                                                     "sample_distance", "start_vel", "end_vel", "start_acc", "end_acc")})
    tol = COEFF_TOL if order < 5 else 1e-6      # the dense reference arithmetic itself degrades at order 5
    for b, p in enumerate(paths):
        s_o, info = mo.generate_trajectory_matrix(p, ocfg)
        sl = res.segment_slice(b)
        assert np.array_equal(res.times[sl], info.Time)
        assert res.iters[b] == info.iters
        assert scaled_coeff_err(res.coeff[sl], info.PolyCoeff.reshape(-1, 3, 2 * order), info.Time) <= tol
        s_g = res.trajectory(b)
        assert s_g.shape == s_o.shape and np.max(np.abs(s_g - s_o)) <= SAMPLE_TOL
        assert abs(res.stats[b, 0] - info.max_climb_rate) <= 1e-6 * max(1.0, info.max_climb_rate)
        assert abs(res.stats[b, 1] - info.min_turn_radius) <= 1e-6 * max(1.0, info.min_turn_radius)


def test_batch_equals_singles_bitwise(tool):
    """Batching must not change any bit: trajectory b of a batch == the same trajectory solved alone."""
    wp, so = workloads.cfg5(B=24, seed=5, ns_min=2, ns_max=40)
    cfg = workloads.synthetic_config(4, "shipped")
    res = tool.generate_batch(cfg, wp, seg_offset=so)
    for b in (0, 7, 23):
        p = wp[so[b] + b: so[b + 1] + b + 1]
        one = tool.generate_batch(cfg, p, ns=p.shape[0] - 1)
        sl = res.segment_slice(b)
        assert np.array_equal(one.coeff, res.coeff[sl]) and np.array_equal(one.samples, res.trajectory(b))
        assert one.iters[0] == res.iters[b] and one.max_dev[0] == res.max_dev[b]


@pytest.mark.parametrize("order,ns,weights", [(4, 16, "shipped"), (4, 16, "plain"), (2, 6, "shipped"), (3, 9, "shipped"),
                                               (5, 4, "shipped"), (4, 1, "shipped"), (4, 2, "plain"), (4, 40, "shipped")])
def test_fused_kernel_equals_generic_path(tool, order, ns, weights):
    """The fused persistent kernel (uniform batches, speculative reweighting) and the per-phase generic kernels
    (sequential reweighting) are the same arithmetic: every output must agree bit for bit."""
    wp = workloads.random_walks(37, ns, seed=order * 100 + ns)
    cfg = workloads.synthetic_config(order, weights)
    cfg.start_vel, cfg.end_acc = (0.5, -0.25, 0.1), (0.0, 0.3, -0.1)
    tool.set_reweight_policy(1)          # generic: one thread per trajectory walks the reweighting loop
    try:
        gen = tool.generate_batch(cfg, wp, ns=ns)
    finally:
        tool.set_reweight_policy(0)      # automatic: fused kernel for uniform batches
    fus = tool.generate_batch(cfg, wp, ns=ns)
    for name in ("times", "coeff", "max_dev", "iters", "vw_final", "best_s", "sample_offset", "samples", "stats", "flags"):
        assert np.array_equal(getattr(gen, name), getattr(fus, name)), name


@pytest.mark.parametrize("order", [2, 3, 4, 5])
def test_speculative_generic_path_equals_sequential_loop(tool, order):
    """CSR batches solve all reweighting iterations at once (k_thomas_spec + k_spec_select); the sequential loop of
    policy 1 is the same arithmetic, so every output must agree bit for bit -- including trajectories that stop early."""
    wp, so = workloads.cfg5(B=96, seed=40 + order, ns_min=1, ns_max=48)
    for weights, pw in (("shipped", None), ("shipped", 3e-4)):       # the larger path weight makes some stop early
        cfg = workloads.synthetic_config(order, weights)
        if pw is not None:
            cfg.path_weight = pw
        tool.set_reweight_policy(1)
        try:
            seq = tool.generate_batch(cfg, wp, seg_offset=so)
        finally:
            tool.set_reweight_policy(0)
        spec = tool.generate_batch(cfg, wp, seg_offset=so)
        for name in ("times", "coeff", "max_dev", "iters", "vw_final", "best_s", "sample_offset", "samples", "stats", "flags"):
            assert np.array_equal(getattr(seq, name), getattr(spec, name)), (name, pw)
        if pw is not None and order <= 4:
            assert len(set(seq.iters.tolist())) > 1                   # the early-stop branch is really exercised


@pytest.mark.parametrize("ragged", [False, True])
def test_pipelined_host_path_equals_single_chunk(tool, ragged):
    """The host-pointer entry point cuts big batches into chunks on two streams; the chunking must not change a bit
    and the CSR sample layout must come out identical (offsets rebased on the host)."""
    if ragged:
        wp, so = workloads.cfg5(B=3000, seed=77, ns_min=1, ns_max=24)
        kw = dict(seg_offset=so)
    else:
        wp, ns = workloads.cfg2(B=3000, ns=8, seed=78)
        kw = dict(ns=ns)
    cfg = workloads.synthetic_config(4, "shipped")
    tool.set_host_chunks(1)
    try:
        one = tool.generate_batch(cfg, wp, **kw)
        tool.set_host_chunks(5)
        five = tool.generate_batch(cfg, wp, **kw)
        cap = int(one.sample_offset[1700]) + 3                         # overflow inside the third chunk
        with pytest.raises(MsnapError) as e:
            tool.generate_batch(cfg, wp, capacity=cap, **kw)
    finally:
        tool.set_host_chunks(0)
    for name in ("times", "coeff", "max_dev", "iters", "vw_final", "best_s", "sample_offset", "samples", "stats", "flags"):
        assert np.array_equal(getattr(one, name), getattr(five, name)), name
    part = e.value.partial
    assert np.array_equal(part.sample_offset, one.sample_offset)
    assert np.array_equal(part.samples, one.samples[:cap])
    assert (part.flags[:1700] & 2).max() == 0 and (part.flags[1701:] & 2).min() == 2


def test_zero_copy_outputs_equal_copied_outputs(tool):
    """With pinned host buffers the kernels store coefficients and samples straight into them; the bytes must equal
    those of the copy path, also when the sample buffer is too small."""
    import torch

    wp, ns = workloads.cfg2(B=600, ns=16, seed=91)
    cfg = workloads.synthetic_config(4, "shipped")
    ref = tool.generate_batch(cfg, wp, ns=ns, stats=False)
    tool.set_zero_copy(True)
    n_seg, cap = 600 * ns, ref.samples.shape[0] + 5
    out = {"coeff": torch.empty((n_seg, 3, 8), dtype=torch.float64).pin_memory().numpy(),
           "samples": torch.full((cap, 3), -7.0, dtype=torch.float64).pin_memory().numpy()}
    try:
        zc = tool.generate_batch(cfg, wp, ns=ns, capacity=cap, out=out, stats=False)
        small = int(ref.sample_offset[300]) + 1
        sm = {"coeff": out["coeff"], "samples": torch.full((cap, 3), -7.0, dtype=torch.float64).pin_memory().numpy()}
        with pytest.raises(MsnapError) as e:
            tool.generate_batch(cfg, wp, ns=ns, capacity=small, out={**sm, "samples": sm["samples"][:small]}, stats=False)
    finally:
        tool.set_zero_copy(False)
    for name in ("times", "coeff", "max_dev", "iters", "vw_final", "best_s", "sample_offset", "samples", "flags"):
        assert np.array_equal(getattr(ref, name), getattr(zc, name)), name
    assert np.all(out["samples"][ref.samples.shape[0]:] == -7.0)               # nothing written past the last row
    assert np.array_equal(e.value.partial.samples, ref.samples[:small])
    assert np.all(sm["samples"][small:] == -7.0)


def test_misaligned_device_coefficient_buffer_is_rejected(tool):
    import torch

    wp, ns = workloads.cfg2(B=4, ns=4)
    cfg = workloads.synthetic_config(4, "plain")
    dev = torch.device("cuda", 0)
    d_wp = torch.from_numpy(wp).to(dev)
    off = torch.empty(5, dtype=torch.int64, device=dev)
    samples = torch.empty((4000, 3), dtype=torch.float64, device=dev)
    raw = torch.empty(4 * ns * 24 + 1, dtype=torch.float64, device=dev)
    with pytest.raises(MsnapError):
        tool.generate_batch_dev(cfg, d_wp, off, samples, ns=ns, coeff=raw[1:])     # 8-byte aligned only
    tool.generate_batch_dev(cfg, d_wp, off, samples, ns=ns, coeff=raw[:-1])
    tool.synchronize()
    assert int(off[-1]) > 8
