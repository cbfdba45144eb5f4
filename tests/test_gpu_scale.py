"""Full-size GPU checks at BASELINE.json's configuration sizes, through size-independent properties: waypoint
interpolation, C^(o-1) continuity, boundary conditions, sampler invariants, determinism, and spot parity against the
oracle port on a seeded subsample.  (pytest -m gpu)"""
import numpy as np
import pytest

from cs_pathplan_b200 import workloads
from helpers import COEFF_TOL, SAMPLE_TOL, scaled_coeff_err

pytestmark = pytest.mark.gpu


def polyval_hi(c, t):
    """c [..., m] highest power first, t [...] -> values."""
    v = np.zeros(c.shape[:-1])
    for i in range(c.shape[-1]):
        v = v * t + c[..., i]
    return v


def polyder_hi(c, r):
    m = c.shape[-1]
    for _ in range(r):
        p = np.arange(c.shape[-1] - 1, 0, -1)
        c = c[..., :-1] * p
    return c


def check_properties(res, wp, seg_offset, order, cfg, pos_tol=2e-8):
    B = seg_offset.shape[0] - 1
    ns = np.diff(seg_offset)
    first_pt = seg_offset[:-1] + np.arange(B)
    seg_traj = np.repeat(np.arange(B), ns)
    seg_pt = np.arange(seg_offset[-1]) + seg_traj                  # waypoint index of each segment's start
    T = res.times
    assert np.all(res.flags == 0)
    assert np.all(T >= cfg.min_time_s)
    c = res.coeff
    # interpolation: p_k(0) = P_k, p_k(T_k) = P_{k+1}
    assert np.max(np.abs(c[:, :, -1] - wp[seg_pt])) == 0.0        # constant term IS the waypoint
    endv = polyval_hi(c, T[:, None])
    assert np.max(np.abs(endv - wp[seg_pt + 1])) <= pos_tol
    # C^(o-1) continuity at interior waypoints and boundary derivatives.  A position-scaled coefficient error eps
    # (the parity metric, bar 1e-8) shows up in the r-th derivative as eps * posmag * (m-1)!/(m-1-r)! / T^r, so that
    # is the yardstick (tolerance 1e-9 of it).
    m = c.shape[-1]
    posmag = np.max(np.abs(c) * T[:, None, None] ** np.arange(m - 1, -1, -1), axis=2)     # [seg, axis]
    interior = np.nonzero(np.arange(seg_offset[-1]) + 1 < seg_offset[seg_traj + 1])[0]
    last_seg = seg_offset[1:] - 1
    for r in range(1, order):
        fall = np.prod(np.arange(m - r, m, dtype=float))
        yard = fall * np.maximum(posmag / T[:, None] ** r, 1.0)
        left = polyval_hi(polyder_hi(c[interior], r), T[interior, None])
        right = polyder_hi(c[interior + 1], r)[..., -1]
        assert np.max(np.abs(left - right) / np.maximum(yard[interior], yard[interior + 1])) <= 1e-9
        if r < 3:   # start/end velocity and acceleration are zero in the synthetic configs
            assert np.max(np.abs(polyder_hi(c[seg_offset[:-1]], r)[..., -1])) <= 1e-9
            endd = polyval_hi(polyder_hi(c[last_seg], r), T[last_seg, None])
            assert np.max(np.abs(endd) / yard[last_seg]) <= 1e-9
    # sampler: first row = first waypoint, last row = last waypoint, spacing >= sample_distance except at seams
    so = res.sample_offset
    assert np.all(np.diff(so) >= 2) and so[-1] == res.samples.shape[0]
    assert np.max(np.abs(res.samples[so[:-1]] - wp[first_pt])) <= pos_tol
    # the end point is appended unless the last accepted candidate already lies within 1e-6 m of it (ms.cpp:159)
    assert np.max(np.abs(res.samples[so[1:] - 1] - wp[first_pt + ns])) <= 1e-6 + pos_tol
    assert np.all(res.iters >= 0) and np.all(res.iters <= 10)
    assert np.all((res.max_dev <= 0.2) | (res.iters == 10))        # the loop's exit condition (ms.cpp:82)


def spot_check_vs_port(res, wp, seg_offset, cfg, picks):
    from oracle import msnap_oracle as mo

    ocfg = mo.MinimumSnapConfig(order=cfg.order, path_weight=cfg.path_weight, vel_zero_weight=cfg.vel_zero_weight,
                                V_avg=cfg.V_avg, min_time_s=cfg.min_time_s, sample_distance=cfg.sample_distance)
    for b in picks:
        p = wp[seg_offset[b] + b: seg_offset[b + 1] + b + 1]
        s_o, info = mo.generate_trajectory_matrix(p, ocfg)
        sl = res.segment_slice(b)
        assert np.array_equal(res.times[sl], info.Time) and res.iters[b] == info.iters
        assert scaled_coeff_err(res.coeff[sl], info.PolyCoeff.reshape(-1, 3, 2 * cfg.order), info.Time) <= COEFF_TOL
        s_g = res.trajectory(b)
        assert s_g.shape == s_o.shape and np.max(np.abs(s_g - s_o)) <= SAMPLE_TOL


@pytest.mark.parametrize("weights", ["plain", "shipped"])
def test_cfg2_full_size(tool, weights):
    wp, ns = workloads.cfg2()
    cfg = workloads.synthetic_config(4, weights)
    res = tool.generate_batch(cfg, wp, ns=ns)
    so = np.arange(4097, dtype=np.int64) * ns
    check_properties(res, wp, so, 4, cfg)
    spot_check_vs_port(res, wp, so, cfg, [0, 1, 2047, 4095])
    again = tool.generate_batch(cfg, wp, ns=ns)                     # bitwise deterministic
    assert np.array_equal(again.coeff, res.coeff) and np.array_equal(again.samples, res.samples)


def test_cfg3_shard_size(tool):
    """cfg3 is 2^20 x 8; one eighth of it (a single GPU's shard at 8 GPUs) is checked here in full, and the shard
    results must equal the same rows of a differently-split batch (sharding is invisible in the results)."""
    wp, ns = workloads.cfg3(B=1 << 17)
    cfg = workloads.synthetic_config(4, "shipped")
    res = tool.generate_batch(cfg, wp, ns=ns)
    so = np.arange((1 << 17) + 1, dtype=np.int64) * ns
    check_properties(res, wp, so, 4, cfg)
    spot_check_vs_port(res, wp, so, cfg, [0, 65535, 131071])
    half = tool.generate_batch(cfg, wp[(1 << 16) * (ns + 1):], ns=ns)
    assert np.array_equal(half.coeff, res.coeff[(1 << 16) * ns:])
    assert np.array_equal(half.samples, res.samples[res.sample_offset[1 << 16]:])


def spot_check_vs_structured(res, wp, seg_offset, cfg, picks):
    """Parity for chains the dense reference arithmetic cannot reach (512 segments = a 4 096 x 4 096 dense inverse, 19 times
    per solve): Oracle B (oracle/msnap_structured.py) replays the reference's loop in double for the discrete decisions
    and in 40-digit arithmetic for the exact coefficients."""
    import mpmath

    from helpers import decisions_equivalent
    from oracle import msnap_structured as st

    mpmath.mp.dps = 40
    Vel = np.array([cfg.start_vel, cfg.end_vel], dtype=float)
    Acc = np.array([cfg.start_acc, cfg.end_acc], dtype=float)
    for b in picks:
        p = wp[seg_offset[b] + b: seg_offset[b + 1] + b + 1]
        sl = res.segment_slice(b)
        d = p[1:] - p[:-1]
        T = np.maximum(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]) / cfg.V_avg, cfg.min_time_s)
        assert np.array_equal(res.times[sl], T)                                     # bit-exact time allocation
        out = st.reweighted_structured(cfg.order, p, Vel, Acc, T, cfg.path_weight, cfg.vel_zero_weight)
        assert out["iters"] == res.iters[b] and out["vw_final"] == res.vw_final[b]
        assert decisions_equivalent(res.best_s[sl], out["best_s"], np.array(out["dist2"], dtype=float))
        truth = st.solve_structured(cfg.order, p, Vel, Acc, T, cfg.path_weight, float(res.vw_final[b]), ctx=mpmath.mp,
                                    best_s=[int(v) for v in res.best_s[sl]])
        tc = np.array([[[float(v) for v in ax] for ax in seg] for seg in truth["coeff"]])
        assert scaled_coeff_err(res.coeff[sl], tc, T) <= COEFF_TOL
        assert abs(res.max_dev[b] - float(truth["max_dev"])) <= 1e-8


def test_cfg4_long_chains(tool):
    """cfg4 at full size: 1 024 boustrophedon patrols x 512 segments."""
    wp, ns = workloads.cfg4()
    cfg = workloads.synthetic_config(4, "shipped")
    res = tool.generate_batch(cfg, wp, ns=ns)
    so = np.arange(1025, dtype=np.int64) * ns
    check_properties(res, wp, so, 4, cfg, pos_tol=1e-7)
    spot_check_vs_structured(res, wp, so, cfg, [0, 1023])
    sub = tool.generate_batch(cfg, wp[700 * (ns + 1): 703 * (ns + 1)], ns=ns)        # batch shape does not matter
    assert np.array_equal(sub.coeff, res.coeff[700 * ns: 703 * ns])
    assert np.array_equal(sub.samples, res.samples[res.sample_offset[700]: res.sample_offset[703]])


def test_very_long_uniform_trajectories(tool):
    """3 000 segments per trajectory: too long for the fused kernel's and the single-launch sampler's shared-memory tiles,
    so the per-pass kernels with the speculative reweighting lanes and the lane-pair solves run; a ragged batch holding
    the same trajectories must give the same bits."""
    wp, ns = workloads.cfg4(B=3, ns=3000, seed=9)
    cfg = workloads.synthetic_config(4, "shipped")
    res = tool.generate_batch(cfg, wp, ns=ns)
    so = np.arange(4, dtype=np.int64) * ns
    check_properties(res, wp, so, 4, cfg, pos_tol=1e-7)
    rag = tool.generate_batch(cfg, wp, seg_offset=so)
    assert np.array_equal(rag.coeff, res.coeff) and np.array_equal(rag.samples, res.samples)
    assert np.array_equal(rag.iters, res.iters) and np.array_equal(rag.max_dev, res.max_dev)


def test_cfg5_mixed_lengths_dense_sampling(tool):
    wp, so = workloads.cfg5(B=8192)
    cfg = workloads.synthetic_config(4, "plain", sample_distance=0.0)
    res = tool.generate_batch(cfg, wp, seg_offset=so)
    check_properties(res, wp, so, 4, cfg)
    # sample_distance 0 accepts every 10 Hz candidate: count = 1 + sum_k n_cand(T_k) (+1 end point unless the last
    # candidate already sits on it)
    T = res.times
    dt = np.minimum(0.1, T / 10.0)
    approx = np.floor((T + 1e-12) / dt + 1e-9)
    per_traj = np.add.reduceat(approx, so[:-1])
    counts = np.diff(res.sample_offset)
    assert np.all(np.abs(counts - (per_traj + 1)) <= np.diff(so) + 1)
    spot_check_vs_port(res, wp, so, cfg, [0, 1000, 2047, 8191])
