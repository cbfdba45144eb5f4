"""GPU parity tests of the batched WGS84 <-> ENU maps, through the C ABI
(msnap_wgs84_to_enu_* / msnap_enu_to_wgs84_* / msnap_set_sample_frame in include/msnap.h).  The checker is
oracle/geo_port.c, itself pinned bit for bit on the reference's recorded run (readme.md:11-28).

Tolerances.  The maps are floating point and the only arithmetic that differs from the reference's is inside
sin / cos / atan2 (CUDA libm vs glibc, <= 2 ulp each), so the bar is the north star's position bar, 1e-6 m:
    ENU metres                : |d| <= 1e-6 m     (asserted at POS_TOL = 1e-7)
    longitude / latitude      : |d| <= 9e-12 deg  (= 1e-6 m on the ground; asserted at ANG_TOL = 1e-12 deg ~ 1e-7 m)
    altitude                  : |d| <= 1e-6 m     (asserted at POS_TOL + ALT_COND / cos(lat), see below)
The number of fixed-point steps (cpp:939-949) is a discrete decision taken at a 1e-12 rad threshold, where the
reference's own `lat_new - lat` carries 1e-4 relative rounding noise.  Where an ulp moves the decision, the latitude
moves by <= 7e-15 rad (the iteration contracts by e^2 = 0.0067 per step) = 4e-8 m on the ground, but the altitude,
which the reference forms as p / cos(lat) - N (cpp:959), moves by R * 7e-15 / cos(lat) = 4.5e-8 m / cos(lat):
6e-8 m at 40 degrees, 1e-6 m at 87.4 degrees.  That sensitivity is the reference formula's (any other libm or compiler
moves its result the same way), so the altitude bar carries the 1 / cos(lat) term."""
import os

import numpy as np
import pytest
import torch

from cs_pathplan_b200 import shipped_config, workloads
from cs_pathplan_b200._lib import MsnapError
from oracle import geo

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
POS_TOL = 1e-7    # metres
ANG_TOL = 1e-12   # degrees
ALT_COND = 1e-7   # metres * cos(lat): one fixed-point step more or less, seen through p / cos(lat) - N


@pytest.fixture(params=["directions", "exact_trig"])
def geo_tool(tool, request):
    """Both execution forms of ecefToWGS84's iteration (msnap_set_geo_exact_trig) must meet the same bars."""
    tool.set_geo_exact_trig(request.param == "exact_trig")
    yield tool
    tool.set_geo_exact_trig(False)


def lla_close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(a - b)
    coslat = np.maximum(np.cos(np.radians(b[:, 1])), 1e-9)
    # a metre on the ground is 1 / cos(lat) times more degrees of longitude than of latitude
    return ((d[:, 0] * coslat).max() <= ANG_TOL and d[:, 1].max() <= ANG_TOL
            and (d[:, 2] - ALT_COND / coslat).max() <= POS_TOL)


def test_readme_waypoints_both_ways(geo_tool):
    tool = geo_tool
    """The reference's only recorded numbers for this path (readme.md:11-28)."""
    enu = tool.wgs84ToENU_Batch(geo.README_WGS84, geo.README_ORIGIN)
    assert np.abs(enu - geo.README_ENU).max() <= POS_TOL
    lla = tool.enuToWGS84_Batch(geo.README_ENU, geo.README_ORIGIN)
    assert lla_close(lla, geo.README_WGS84_BACK)
    # one-point methods of the reference class
    assert np.array_equal(tool.wgs84ToENU(geo.README_WGS84[3], geo.README_ORIGIN), enu[3])
    assert np.array_equal(tool.enuToWGS84(geo.README_ENU[3], geo.README_ORIGIN), lla[3])


def test_committed_fixture(geo_tool):
    tool = geo_tool
    z = np.load(os.path.join(HERE, "golden", "geo_golden.npz"))
    same_steps = total = 0
    for r, ref in enumerate(z["refs"]):
        lla = tool.enuToWGS84_Batch(z["enu"][r], ref)
        assert lla_close(lla, z["lla"][r]), r
        assert np.abs(tool.wgs84ToENU_Batch(z["lla"][r], ref) - z["enu_back"][r]).max() <= POS_TOL, r
        d_e = torch.from_numpy(z["enu"][r]).cuda()
        d_l = torch.empty_like(d_e)
        d_s = torch.zeros(d_e.shape[0], dtype=torch.int32, device="cuda")
        tool.enu_to_wgs84_dev(ref, d_e, d_l, steps_out=d_s)
        tool.synchronize()
        assert np.array_equal(d_l.cpu().numpy(), lla)                      # device rows == host rows, bitwise
        st = d_s.cpu().numpy()
        assert np.all(np.abs(st - z["steps"][r]) <= 1)
        same_steps += int(np.sum(st == z["steps"][r]))
        total += st.size
    assert same_steps >= 0.99 * total                                         # the stopping decision itself agrees


def test_seeded_points_against_the_port(geo_tool):
    tool = geo_tool
    rng = np.random.default_rng(99)
    ref = np.array([116.39, 39.91, 43.5])
    n = 200_003                                                               # ragged: not a multiple of 32 or 256
    enu = np.column_stack([rng.normal(0, 5e4, n), rng.normal(0, 5e4, n), rng.uniform(-100, 9000, n)])
    lla = tool.enuToWGS84_Batch(enu, ref)
    assert lla_close(lla, geo.enu_to_wgs84_batch(enu, ref, threads=8))
    back = tool.wgs84ToENU_Batch(lla, ref)
    assert np.abs(back - geo.wgs84_to_enu_batch(lla, ref, threads=8)).max() <= POS_TOL
    assert np.abs(back - enu).max() <= 2e-7                                   # round trip (1e-12 rad stopping rule)
    # a batch is the same as its points one by one, and in place is the same as out of place
    for i in (0, 31, 32, 100_000, n - 1):
        assert np.array_equal(tool.enuToWGS84(enu[i], ref), lla[i])
    d = torch.from_numpy(enu).cuda()
    tool.enu_to_wgs84_dev(ref, d, d)
    tool.synchronize()
    assert np.array_equal(d.cpu().numpy(), lla)
    tool.wgs84_to_enu_dev(ref, d, d)
    tool.synchronize()
    assert np.array_equal(d.cpu().numpy(), back)


def test_empty_and_invalid(geo_tool):
    tool = geo_tool
    assert tool.enuToWGS84_Batch(np.zeros((0, 3)), [0, 0, 0]).shape == (0, 3)
    assert tool.wgs84ToENU_Batch(np.zeros((0, 3)), [0, 0, 0]).shape == (0, 3)
    with pytest.raises(MsnapError):
        tool.enuToWGS84_Batch(np.zeros((4, 3)), [np.nan, 0, 0])
    with pytest.raises(ValueError):
        tool.enuToWGS84_Batch(np.zeros((4, 2)), [0, 0, 0])
    with pytest.raises(MsnapError):
        tool.set_sample_frame("wgs84", None)
    # exactly on the polar axis the reference's latitude is NaN (0 * inf in cpp:944) and its altitude comes from the
    # pole branch (cpp:956-957): same here
    ref = np.array([0.0, 90.0, 0.0])
    x = geo.wgs84_to_ecef(ref)[0]
    out = tool.enuToWGS84_Batch(np.array([[0.0, x, 25.0]]), ref)[0]
    exp = geo.enu_to_wgs84_batch(np.array([[0.0, x, 25.0]]), ref)[0]
    assert np.isnan(out[1]) == np.isnan(exp[1]) and abs(out[2] - exp[2]) <= POS_TOL


def test_direction_form_equals_trig_form(tool):
    """The default (direction-vector) iteration against the statement-by-statement one: same steps, rounding-level
    differences in the results."""
    rng = np.random.default_rng(3)
    n = 1 << 20
    for ref in ([109.56, 40.87, 0.0], [10.0, -89.2, 2800.0], [-45.0, 0.01, 0.0]):
        ref = np.array(ref)
        enu = np.column_stack([rng.normal(0, 1e5, n), rng.normal(0, 1e5, n), rng.uniform(-1000, 20000, n)])
        d = torch.from_numpy(enu).cuda()
        out = {}
        for trig in (False, True):
            tool.set_geo_exact_trig(trig)
            lla = torch.empty_like(d)
            st = torch.zeros(n, dtype=torch.int32, device="cuda")
            tool.enu_to_wgs84_dev(ref, d, lla, steps_out=st)
            tool.synchronize()
            out[trig] = (lla.cpu().numpy(), st.cpu().numpy())
        tool.set_geo_exact_trig(False)
        fin = np.isfinite(out[True][0]).all(axis=1)
        assert fin.mean() > 0.999999 and np.array_equal(fin, np.isfinite(out[False][0]).all(axis=1))
        assert lla_close(out[False][0][fin], out[True][0][fin])
        assert np.abs(out[False][1] - out[True][1]).max() <= 1
        assert np.mean(out[False][1] == out[True][1]) >= 0.99


def test_sampler_rows_leave_as_wgs84(tool):
    """msnap_set_sample_frame(1): generate's rows == enuToWGS84_Batch of the ENU rows (getPlan, cpp:3699), bitwise the
    standalone kernel's, statistics and everything else unchanged; uniform (fused) and ragged (generic) batches, host
    chunks included."""
    origin = geo.README_ORIGIN
    try:
        for kind in ("uniform", "ragged", "chunked"):
            if kind == "ragged":
                wp, so = workloads.cfg5(B=300, seed=5)
                kw = dict(seg_offset=so)
            else:
                B = 20_000 if kind == "chunked" else 500
                wp, ns = workloads.cfg2(B=B, ns=16)
                kw = dict(ns=ns)
            cfg = workloads.synthetic_config(4, "shipped")
            tool.set_sample_frame("enu")
            a = tool.generate_batch(cfg, wp, **kw)
            tool.set_sample_frame("wgs84", origin)
            b = tool.generate_batch(cfg, wp, **kw)
            assert np.array_equal(a.sample_offset, b.sample_offset) and np.array_equal(a.coeff, b.coeff)
            assert np.array_equal(a.stats, b.stats) and np.array_equal(a.iters, b.iters)
            exp = tool.enuToWGS84_Batch(a.samples, origin)
            assert np.array_equal(b.samples, exp), kind
            assert lla_close(b.samples[:5000], geo.enu_to_wgs84_batch(a.samples[:5000], origin))
        # the reference's own case, through the reference-shaped method
        tool.set_sample_frame("wgs84", origin)
        s = tool.GenerateTrajectoryMatrix(workloads.UAV31_0_ENU, shipped_config(), 300.0, 30.0)
        tool.set_sample_frame("enu")
        e = tool.GenerateTrajectoryMatrix(workloads.UAV31_0_ENU, shipped_config(), 300.0, 30.0)
        assert s.shape == e.shape == (168, 3)
        assert lla_close(s, geo.enu_to_wgs84_batch(e, origin))
    finally:
        tool.set_sample_frame("enu")


def test_full_size_round_trip_properties(tool):
    """16 M points (the size of a cfg3 shard's sample output): size-independent properties only."""
    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(5)
    enu = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    enu[:, :2] = torch.randn((n, 2), generator=g, dtype=torch.float64, device="cuda") * 2.0e4
    enu[:, 2] = torch.rand(n, generator=g, dtype=torch.float64, device="cuda") * 5000.0
    ref = np.array([109.56059880227296, 40.86719901015758, 0.0])
    lla = torch.empty_like(enu)
    back = torch.empty_like(enu)
    tool.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        tool.enu_to_wgs84_dev(ref, enu, lla)
        tool.wgs84_to_enu_dev(ref, lla, back)
        torch.cuda.synchronize()
    finally:
        tool.set_stream(None)
    assert float((back - enu).abs().max()) <= 2e-7
    assert bool(torch.isfinite(lla).all())
    # altitude ~ up + curvature drop: d^2 / (2 R) within a few percent for d << R
    d2 = enu[:, 0] ** 2 + enu[:, 1] ** 2
    drop = lla[:, 2] - enu[:, 2]
    assert float((drop - d2 / (2 * 6.37e6)).abs().max()) <= 0.02 * float((d2 / (2 * 6.37e6)).max()) + 1e-3
    # spot check against the port on a strided subsample
    idx = torch.arange(0, n, 4099, device="cuda")
    assert lla_close(lla[idx].cpu().numpy(), geo.enu_to_wgs84_batch(enu[idx].cpu().numpy(), ref, threads=8))


def test_wgs84_waypoints_in_wgs84_rows_out(tool):
    """getPlan's leader chain on the device (cpp:2640 -> 3684 -> 3699): WGS84 waypoints -> ENU -> minimum snap ->
    sampled ENU -> WGS84 rows.  (1) the reference's own case against the oracle chain, (2) a batch: bitwise the same as
    the three calls made one after the other."""
    from oracle import msnap_oracle as mo

    origin = geo.README_ORIGIN
    try:
        tool.set_waypoint_frame("wgs84", origin)
        tool.set_sample_frame("wgs84", origin)
        s = tool.GenerateTrajectoryMatrix(geo.README_WGS84, shipped_config(), 300.0, 30.0)
        enu_wp = geo.wgs84_to_enu_batch(geo.README_WGS84, origin)
        ocfg = mo.MinimumSnapConfig(order=2, path_weight=1e-7, vel_zero_weight=0.01, V_avg=200.0, min_time_s=1.0,
                                    sample_distance=300.0)
        s_o, _ = mo.generate_trajectory_matrix(enu_wp, ocfg, sample_distance_override=300.0, v_avg_override=30.0)
        assert s.shape == s_o.shape == (168, 3)
        assert lla_close(s, geo.enu_to_wgs84_batch(s_o, origin))
        # batch
        wp_enu, ns = workloads.cfg2(B=700, ns=16)
        wp_enu = wp_enu * np.array([30.0, 30.0, 1.0]) + np.array([0.0, 0.0, 1500.0])   # a few km apart, 1.5 km up
        wp_lla = geo.enu_to_wgs84_batch(wp_enu, origin)
        cfg = workloads.synthetic_config(4, "shipped", sample_distance=25.0)
        cfg.V_avg = 150.0
        both = tool.generate_batch(cfg, wp_lla, ns=ns)
        tool.set_waypoint_frame("enu")
        tool.set_sample_frame("enu")
        step1 = tool.wgs84ToENU_Batch(wp_lla, origin)
        step2 = tool.generate_batch(cfg, step1, ns=ns)
        step3 = tool.enuToWGS84_Batch(step2.samples, origin)
        assert np.array_equal(both.sample_offset, step2.sample_offset) and np.array_equal(both.coeff, step2.coeff)
        assert np.array_equal(both.times, step2.times) and np.array_equal(both.samples, step3)
        assert np.abs(step1 - wp_enu).max() <= 2e-7
    finally:
        tool.set_waypoint_frame("enu")
        tool.set_sample_frame("enu")
