"""Robustness of the device entry points against inputs that would corrupt memory or hang a launch (round-1 advisor
findings): a truncated generate chained into the altitude stage, non-finite / absurd waypoints, parameters the sampler's
candidate loop cannot terminate with, and the reference's failure semantics of optimizeSegmentAltitudeENU
(uavPathPlanning.cpp:1342-1344, 1356).  (pytest -m gpu)"""
import numpy as np
import pytest
import torch

from cs_pathplan_b200 import MinimumSnapConfig, shipped_altitude_params, workloads
from cs_pathplan_b200._lib import ERR_INVALID_ARG, MsnapError

pytestmark = pytest.mark.gpu


def test_truncated_generate_chained_into_altitude_stays_inside_the_buffers(tool):
    """msnap_generate_batch_dev with too little sample capacity still writes the EXACT row layout; the altitude stage fed
    with that layout must not touch rows beyond the capacity: trajectories that do not fit are skipped and flagged."""
    B, ns = 64, 16
    wp, _ = workloads.cfg2(B=B, ns=ns)
    cfg = workloads.synthetic_config(4, "plain", sample_distance=2.0)
    full = tool.generate_batch(cfg, wp, ns=ns)
    n_full = int(full.sample_offset[-1])
    cut_b = 40
    cap = int(full.sample_offset[cut_b]) + 5                    # trajectory 40 fits only partly, 41.. not at all
    dev = torch.device("cuda")
    guard = 4096                                                # canary rows behind the declared capacity
    d_rows = torch.full((cap + guard, 3), -777.0, dtype=torch.float64, device=dev)
    d_elev = torch.full((cap + guard,), 1.0e5, dtype=torch.float64, device=dev)  # a terrain that would lift every row
    d_off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    d_flags = torch.zeros(B, dtype=torch.int32, device=dev)
    d_aflags = torch.zeros(B, dtype=torch.int32, device=dev)
    d_z1 = torch.full((cap + guard,), -777.0, dtype=torch.float64, device=dev)
    tool.generate_batch_dev(cfg, torch.from_numpy(wp).to(dev), d_off, d_rows[:cap], ns=ns, flags=d_flags)
    tool.synchronize()
    assert int(d_off[-1]) == n_full > cap and np.array_equal(d_off.cpu().numpy(), full.sample_offset)
    before = d_rows.cpu().numpy().copy()
    assert np.all(before[cap:] == -777.0)
    p = shipped_altitude_params()
    tool.altitude_optimize_batch_dev(p, d_off, d_rows[:cap], d_elev[:cap], z_pass1=d_z1[:cap], flags=d_aflags)
    tool.synchronize()
    after = d_rows.cpu().numpy()
    assert np.all(after[cap:] == -777.0) and np.all(d_z1.cpu().numpy()[cap:] == -777.0)   # nothing behind the capacity
    fl = d_aflags.cpu().numpy().view(np.uint32)
    assert np.all(fl[:cut_b] == 0) and np.all(fl[cut_b:] == 2)                            # skipped = TRUNCATED
    lo = int(full.sample_offset[cut_b])
    assert np.array_equal(after[lo:cap], before[lo:cap])                                  # the partial trajectory is untouched
    assert np.all(after[:lo, 2] >= 1.0e5)                                                 # the complete ones were optimised
    # and equal to the same trajectories optimised on their own
    ref = tool.altitude_optimize_batch(full.samples[:lo], full.sample_offset[:cut_b + 1], p, np.full(lo, 1.0e5))
    assert np.array_equal(ref, after[:lo])


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("bad", [np.inf, np.nan, 1.0e300, 3.0e8])
def test_one_bad_waypoint_does_not_hang_or_disturb_the_batch(tool, ragged, bad):
    """T = inf / NaN / 1e299 s / 6e7 s in ONE trajectory: the launch returns, that trajectory is flagged and gets no sampled
    candidates, every other trajectory has the bits it has without the bad neighbour."""
    B, ns = 96, 16
    wp, _ = workloads.cfg2(B=B, ns=ns)
    cfg = workloads.synthetic_config(4, "shipped")
    so = np.arange(B + 1, dtype=np.int64) * ns
    kw = dict(seg_offset=so) if ragged else dict(ns=ns)
    good = tool.generate_batch(cfg, wp, **kw)
    wpb = wp.copy()
    b_bad = 37
    wpb[b_bad * (ns + 1) + 5, 0] = bad
    cap = int(good.sample_offset[-1]) + 64
    try:
        res = tool.generate_batch(cfg, wpb, capacity=cap, **kw)
    except MsnapError as e:                        # (sample bound differs: the capacity may not fit; layout still exact)
        res = e.partial
    assert res.flags[b_bad] & 1
    assert np.all(np.delete(res.flags, b_bad) == 0)
    for b in (0, b_bad - 1, b_bad + 1, B - 1):
        assert np.array_equal(res.trajectory(b), good.trajectory(b))
        assert np.array_equal(res.coeff[res.segment_slice(b)], good.coeff[good.segment_slice(b)])
    assert res.sample_offset[b_bad + 1] - res.sample_offset[b_bad] <= good.sample_offset[b_bad + 1] - good.sample_offset[b_bad]


def test_parameters_the_sampler_cannot_terminate_with_are_rejected(tool):
    wp, ns = workloads.cfg2(B=4, ns=4)
    for kw in (dict(min_time_s=0.0), dict(min_time_s=-1.0), dict(V_avg=float("inf")), dict(sample_distance=float("nan")),
               dict(min_time_s=1.0e9)):
        cfg = MinimumSnapConfig(order=3, **kw)
        with pytest.raises(MsnapError) as e:
            tool.generate_batch(cfg, wp, ns=ns, capacity=1000)
        assert e.value.status == ERR_INVALID_ARG
    # the reference's own edge case stays legal: V_avg = 0 means every segment takes min_time_s (ms.cpp:66-70)
    res = tool.generate_batch(MinimumSnapConfig(order=3, V_avg=0.0, min_time_s=2.0), wp, ns=ns)
    assert np.all(res.times == 2.0) and not res.flags.any()


def test_altitude_failure_semantics_follow_the_reference(tool):
    """optimizeSegmentAltitudeENU returns false and leaves the segment untouched when optimizeHeights' factorisation fails
    (cpp:1342-1344, 1671-1675).  The Hessian depends on the parameters and the horizontal geometry only, so the failure is
    provoked through a non-finite follow weight: every pivot is NaN, every trajectory keeps its rows, bit 0 is set."""
    from alt_helpers import sampled_paths
    from cs_pathplan_b200 import AltitudeParams

    rows, off = sampled_paths(12, seed=3, n_min=30, n_max=60)
    elev = np.full(rows.shape[0], 900.0)
    good, _, _, fl0 = tool.altitude_optimize_batch(rows, off, shipped_altitude_params(), elev, return_info=True)
    assert not fl0.any() and not np.array_equal(good, rows)
    bad = AltitudeParams(lambda_smooth=1.0, lambda_follow=float("nan"), max_climb_rate=0.3, safe_distance=10.0)
    for policy in (0, 1, 2):
        tool.set_altitude_policy(policy)
        out, _, _, fl = tool.altitude_optimize_batch(rows, off, bad, elev, return_info=True)
        assert np.all(fl & 1) and np.array_equal(out, rows)          # untouched
    tool.set_altitude_policy(2)


@pytest.mark.parametrize("sd,v", [(300.0, 30.0), (5.0, 30.0), (40.0, 200.0), (300.0, 8.0)])   # (v = 8: T up to 4 000 s, beyond the time table)
def test_long_segments_warp_cooperative_sampler_equals_sequential_loop(tool, sd, v):
    """Segments with more than 128 candidates (T > 12.8 s: the reference's shipped mission has 7 000 per segment) are walked by
    a whole warp in the single-launch sampler.  Rows must equal, bit for bit, those of the per-pass kernels (policy 1), where
    one lane walks the candidates in the reference's order -- on a batch that mixes sub-second, ordinary and very long
    segments -- and the compiled reference on a few of them."""
    from cs_pathplan_b200 import shipped_config
    from oracle import ref

    rng = np.random.default_rng(17)
    B, ns = 300, 6
    legs = rng.choice([3.0, 40.0, 300.0, 2500.0, 22000.0], size=(B, ns, 1)) * rng.uniform(0.5, 1.5, (B, ns, 1))
    d = rng.normal(size=(B, ns, 3))
    d[..., 2] *= 0.05
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    wp = np.concatenate([np.zeros((B, 1, 3)), np.cumsum(d * legs, axis=1)], axis=1).reshape(-1, 3)
    cfg = shipped_config()
    fast = tool.generate_batch(cfg, wp, ns=ns, sample_distance_override=sd, v_avg_override=v)
    assert not fast.flags.any() and (fast.times > 12.8).sum() > B and (fast.times < 1.0 + 1e-9).sum() > 0
    tool.set_reweight_policy(1)
    try:
        slow = tool.generate_batch(cfg, wp, ns=ns, sample_distance_override=sd, v_avg_override=v)
    finally:
        tool.set_reweight_policy(0)
    assert np.array_equal(fast.sample_offset, slow.sample_offset) and np.array_equal(fast.samples, slow.samples)
    assert np.array_equal(fast.stats, slow.stats)
    rc = ref.shipped_config()
    for b in (0, 7, 150, B - 1):
        want = ref.generate(wp[b * (ns + 1):(b + 1) * (ns + 1)], rc, sd, v)
        got = fast.trajectory(b)
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-6, b
    # a capacity that cuts through the rows of a long segment: exact layout, rows that fit are right, flagged
    cap = int(fast.sample_offset[B // 2]) + 3
    with pytest.raises(MsnapError) as e:
        tool.generate_batch(cfg, wp, ns=ns, sample_distance_override=sd, v_avg_override=v, capacity=cap)
    part = e.value.partial
    assert np.array_equal(part.sample_offset, fast.sample_offset) and np.array_equal(part.samples, fast.samples[:cap])
    assert (part.flags[B // 2 + 1:] & 2).all() and not (part.flags[: B // 2] & 2).any()
