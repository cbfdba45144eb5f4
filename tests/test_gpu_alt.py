"""GPU parity tests of the batched altitude optimisation through the C ABI
(msnap_altitude_optimize_batch_* / msnap_cost_map_lookup_dev in include/msnap.h) against oracle/alt_oracle.py.

Bars: optimised heights within 1e-6 m (the north star's position bar) of the port; identical discrete decisions: the
number of solves the active-set loop takes, and -- through the heights -- which rows it pinned.  The port is itself
within 1e-6 m of a 50-digit solve of the same systems (tests/test_alt_oracle.py); the reference's own solver
(Eigen::SimplicialLDLT) is not available in this image, so parity with the reference's bits is unpinned (DESIGN.md
section 10)."""
import numpy as np
import pytest
import torch

from alt_helpers import lookup, sampled_paths, terrain_grid
from cs_pathplan_b200 import AltitudeParams, shipped_altitude_params, workloads
from cs_pathplan_b200._lib import MsnapError
from oracle import alt_oracle as ao
from oracle import geo

pytestmark = pytest.mark.gpu
Z_TOL = 1e-6


POLICIES = {"lane_pairs": 0, "one_lane": 1, "partitioned": 2}


@pytest.fixture(params=list(POLICIES))
def alt_tool(tool, request):
    """All execution forms of the banded solves (msnap_set_altitude_policy) must meet the same bars."""
    tool.set_altitude_policy(POLICIES[request.param])
    yield tool
    tool.set_altitude_policy(2)


def oracle_params(p):
    return ao.AltitudeParams(p.lambda_smooth, p.lambda_follow, p.max_climb_rate, p.uav_R, p.safe_distance)


@pytest.mark.parametrize("params", [shipped_altitude_params(), AltitudeParams(),
                                    AltitudeParams(lambda_smooth=0.0, lambda_follow=2.0, max_climb_rate=0.5, safe_distance=30.0),
                                    AltitudeParams(lambda_smooth=3.0, lambda_follow=0.5, max_climb_rate=0.0, safe_distance=5.0)],
                         ids=["shipped", "struct_defaults", "no_smoothing", "no_climb_term"])
def test_ragged_batch_against_the_port(alt_tool, params):
    tool = alt_tool
    grid, res, ox, oy = terrain_grid()
    rows, off = sampled_paths(48, seed=21)
    elev = lookup(grid, res, ox, oy, rows)
    assert np.isnan(elev).any() and (~np.isnan(elev)).any()          # some rows lie outside the map
    if params.max_climb_rate <= 0.0:
        # without the climb term, rows outside the map are held by the 1e-8 regularisation alone (cond ~ 1e9: the system
        # extrapolates linearly and any two factorisations differ by ~1e-5 m there); give every row a terrain value
        elev = np.where(np.isnan(elev), 1200.0, elev)
    out, z1, solves, flags = tool.altitude_optimize_batch(rows, off, params, elev, return_info=True)
    assert not flags.any()
    assert np.array_equal(out[:, :2], rows[:, :2])                   # only the up column changes (cpp:1357-1359)
    po = oracle_params(params)
    # pass 2 without the climb term is a pure fourth-difference operator between pinned rows: cond ~ n^4 (1e9 at n = 200),
    # so two backward-stable factorisations agree to ~1e-5 m only; with the climb term (every shipped setting) 1e-6 m holds
    tol2 = Z_TOL if params.max_climb_rate > 0.0 else 2e-5
    for b in range(off.shape[0] - 1):
        sl = slice(int(off[b]), int(off[b + 1]))
        z2_o, z1_o, solves_o, _ = ao.optimize_segment_altitude_enu(rows[sl], po, elev[sl], return_info=True)
        assert np.abs(z1[sl] - z1_o).max() <= Z_TOL, b
        assert solves[b] == solves_o, b
        assert np.abs(out[sl, 2] - z2_o).max() <= tol2, b


def test_reference_shaped_single_call_and_edge_cases(alt_tool):
    tool = alt_tool
    p = shipped_altitude_params()
    seg = np.column_stack([np.arange(50) * 30.0, np.zeros(50), np.full(50, 1000.0)])
    elev = 980.0 + 25.0 * np.sin(np.arange(50) / 5.0)
    out = tool.optimizeSegmentAltitudeENU(seg, p, elev)
    exp = ao.optimize_segment_altitude_enu(seg, oracle_params(p), elev)
    assert np.abs(out[:, 2] - exp).max() <= Z_TOL and np.all(out[:, 2] >= elev + p.safe_distance - 1e-9)
    # no terrain at all: elev = None and elev = NaN are the same thing
    a = tool.optimizeSegmentAltitudeENU(seg, p, None)
    b = tool.optimizeSegmentAltitudeENU(seg, p, np.full(50, np.nan))
    assert np.array_equal(a, b)
    assert np.abs(a[:, 2] - ao.optimize_segment_altitude_enu(seg, oracle_params(p), np.full(50, np.nan))).max() <= Z_TOL
    # n = 1, 2, 3 and an empty trajectory inside a batch
    rows = np.column_stack([np.arange(6) * 40.0, np.zeros(6), np.full(6, 1000.0)])
    off = np.array([0, 1, 1, 3, 6], dtype=np.int64)
    el = np.full(6, 995.0)
    out, z1, solves, flags = tool.altitude_optimize_batch(rows, off, p, el, return_info=True)
    assert solves[1] == 0 and not flags.any()
    for b in (0, 2, 3):
        sl = slice(int(off[b]), int(off[b + 1]))
        assert np.abs(out[sl, 2] - ao.optimize_segment_altitude_enu(rows[sl], oracle_params(p), el[sl])).max() <= Z_TOL
    # empty batch, malformed offsets
    assert tool.altitude_optimize_batch(np.zeros((0, 3)), np.array([0]), p).shape == (0, 3)
    with pytest.raises(ValueError):
        tool.altitude_optimize_batch(rows, np.array([0, 5]), p)
    with pytest.raises(MsnapError):
        tool.altitude_optimize_batch(rows, np.array([0, 4, 2, 6]), p)


def test_batch_equals_singles_bitwise(alt_tool):
    tool = alt_tool
    grid, res, ox, oy = terrain_grid()
    rows, off = sampled_paths(40, seed=5)
    elev = lookup(grid, res, ox, oy, rows)
    p = shipped_altitude_params()
    out = tool.altitude_optimize_batch(rows, off, p, elev)
    for b in (0, 3, 7, 39):
        sl = slice(int(off[b]), int(off[b + 1]))
        assert np.array_equal(tool.optimizeSegmentAltitudeENU(rows[sl], p, elev[sl]), out[sl])


def test_device_chain_sampler_lookup_altitude_wgs84(alt_tool):
    tool = alt_tool
    """getPlan's leader chain after Minisnap_3D on the device (cpp:3684 -> 3712-3729 -> 1535-1573): sampled ENU rows ->
    cost-map lookup -> optimizeSegmentAltitudeENU -> enuToWGS84_Batch, against the oracles run step by step on the host."""
    grid, res, ox, oy = terrain_grid(width=900, height=900, resolution=10.0, origin_x=-4500.0, origin_y=4500.0)
    B, ns = 64, 16
    wp, _ = workloads.cfg2(B=B, ns=ns)
    wp = wp * np.array([8.0, 8.0, 1.0]) + np.array([0.0, 0.0, 1250.0])
    cfg = workloads.synthetic_config(4, "shipped", sample_distance=20.0)
    cfg.V_avg = 40.0
    res_enu = tool.generate_batch(cfg, wp, ns=ns)
    n = res_enu.samples.shape[0]
    dev = torch.device("cuda")
    d_rows = torch.from_numpy(res_enu.samples).to(dev)
    d_off = torch.from_numpy(res_enu.sample_offset).to(dev)
    d_grid = torch.from_numpy(grid).to(dev)
    d_elev = torch.empty(n, dtype=torch.float64, device=dev)
    d_solves = torch.zeros(B, dtype=torch.int32, device=dev)
    p = shipped_altitude_params()
    tool.cost_map_lookup_dev(d_grid, res, ox, oy, d_rows, d_elev, n_rows=d_off[B:])
    tool.altitude_optimize_batch_dev(p, d_off, d_rows, d_elev, solves=d_solves)
    d_lla = torch.empty_like(d_rows)
    tool.enu_to_wgs84_dev(geo.README_ORIGIN, d_rows, d_lla)
    tool.synchronize()
    elev_o = lookup(grid, res, ox, oy, res_enu.samples)
    got_elev = d_elev.cpu().numpy()
    assert np.array_equal(np.isnan(got_elev), np.isnan(elev_o)) and np.array_equal(got_elev[~np.isnan(elev_o)], elev_o[~np.isnan(elev_o)])
    rows_o = res_enu.samples.copy()
    for b in range(B):
        sl = slice(int(res_enu.sample_offset[b]), int(res_enu.sample_offset[b + 1]))
        z2, _, solves_o, _ = ao.optimize_segment_altitude_enu(res_enu.samples[sl], oracle_params(p), elev_o[sl], return_info=True)
        rows_o[sl, 2] = z2
        assert int(d_solves[b]) == solves_o
    assert np.abs(d_rows.cpu().numpy() - rows_o).max() <= Z_TOL
    lla_o = geo.enu_to_wgs84_batch(rows_o, geo.README_ORIGIN)
    d = np.abs(d_lla.cpu().numpy() - lla_o)
    assert d[:, :2].max() <= 1e-11 and d[:, 2].max() <= 2e-6
    # host path == device path, bitwise
    host = tool.altitude_optimize_batch(res_enu.samples, res_enu.sample_offset, p, elev_o)
    assert np.array_equal(host, d_rows.cpu().numpy())


def test_full_size_properties(alt_tool):
    tool = alt_tool
    """cfg2-sized sampler output (4 096 trajectories, ~0.8 M rows): clearance, monotonicity and fixed end points."""
    rng = np.random.default_rng(4)
    B = 4096
    ns = rng.integers(150, 260, B)
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    n = int(off[-1])
    t = np.arange(n) - np.repeat(off[:-1], ns)
    rows = np.column_stack([t * 25.0, np.repeat(rng.uniform(-1e3, 1e3, B), ns), 1300.0 + 40.0 * np.sin(t / 9.0)])
    elev = 1250.0 + 80.0 * np.sin(rows[:, 0] / 400.0 + np.repeat(rng.uniform(0, 6, B), ns))
    p = shipped_altitude_params()
    out, z1, solves, flags = tool.altitude_optimize_batch(rows, off, p, elev, return_info=True)
    assert not flags.any() and solves.min() >= 1 and solves.max() <= 10
    assert np.all(z1 >= elev + p.safe_distance - 1e-9)
    assert np.all(out[:, 2] >= z1)
    first, last = off[:-1], off[1:] - 1
    assert np.abs(out[first, 2] - z1[first]).max() <= 1e-6 and np.abs(out[last, 2] - z1[last]).max() <= 1e-6
    for b in (0, 1234, 4095):
        sl = slice(int(off[b]), int(off[b + 1]))
        z2 = ao.optimize_segment_altitude_enu(rows[sl], oracle_params(p), elev[sl])
        assert np.abs(out[sl, 2] - z2).max() <= Z_TOL


def test_lane_pairs_and_partitions_equal_one_lane(tool):
    """Two-sided elimination and the partitioned (nested-dissection) form against the plain downward recurrence: same
    active-set decisions, heights equal to rounding; lengths around the pairing threshold (8 rows), around the partition
    counts (n / 4 lanes up to 16; 270 rows = the last length a group of 16 lanes holds, 542 the last 32 lanes hold in shared
    memory) and odd / even lengths included."""
    grid, res, ox, oy = terrain_grid()
    rows, off = sampled_paths(200, seed=77, n_min=1, n_max=40)
    rows2, off2 = sampled_paths(60, seed=78, n_min=100, n_max=400)
    rows3, off3 = sampled_paths(24, seed=79, n_min=255, n_max=280)
    rows4, off4 = sampled_paths(12, seed=80, n_min=535, n_max=550)
    rows = np.vstack([rows, rows2, rows3, rows4])
    for o in (off2, off3, off4):
        off = np.concatenate([off, off[-1] + o[1:]])
    elev = lookup(grid, res, ox, oy, rows)
    p = shipped_altitude_params()
    out = {}
    try:
        for pol in (0, 1, 2):
            tool.set_altitude_policy(pol)
            out[pol] = tool.altitude_optimize_batch(rows, off, p, elev, return_info=True)
    finally:
        tool.set_altitude_policy(2)
    for pol in (0, 2):
        assert np.array_equal(out[pol][2], out[1][2]), pol                     # solves of the active-set loop
        assert not out[pol][3].any() and not out[1][3].any()
        assert np.abs(out[pol][1] - out[1][1]).max() <= 1e-6, pol              # pass 1 (rows outside the map are weakly held)
        assert np.abs(out[pol][0][:, 2] - out[1][0][:, 2]).max() <= 1e-6, pol  # final heights


def test_long_trajectories_many_chunks(alt_tool):
    """Thousands of rows per trajectory (dozens of 32-row chunks per sweep, both sides of a pair), lengths chosen around
    chunk boundaries; checker: the O(n) banded-Cholesky variant of the port."""
    tool = alt_tool
    rng = np.random.default_rng(31)
    ns = np.array([3000, 1537, 64, 65, 63, 33, 32, 31, 2049, 1000])
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    n = int(off[-1])
    t = np.arange(n) - np.repeat(off[:-1], ns)
    rows = np.column_stack([t * 25.0 + rng.normal(0, 2.0, n), np.repeat(rng.uniform(-500, 500, len(ns)), ns) + rng.normal(0, 2.0, n),
                            1300.0 + 60.0 * np.sin(t / 37.0)])
    elev = 1260.0 + 70.0 * np.sin(rows[:, 0] / 900.0) + 25.0 * np.sin(rows[:, 0] / 130.0)
    p = shipped_altitude_params()
    out, z1, solves, flags = tool.altitude_optimize_batch(rows, off, p, elev, return_info=True)
    assert not flags.any()
    for b in range(len(ns)):
        sl = slice(int(off[b]), int(off[b + 1]))
        z2_o, z1_o, solves_o, _ = ao.optimize_segment_altitude_enu_banded(rows[sl], oracle_params(p), elev[sl], return_info=True)
        assert np.abs(z1[sl] - z1_o).max() <= Z_TOL, b
        assert solves[b] == solves_o, b
        assert np.abs(out[sl, 2] - z2_o).max() <= Z_TOL, b


def test_partitioned_staging_paths_agree(tool):
    """k_alt_part stages a trajectory's rows by bulk copies (16-byte units) when the caller's arrays are 16-byte aligned
    and by per-lane 8-byte copies otherwise; windows that would reach beyond the arrays fetch their last odd element by an
    ordinary load.  Same rows through device buffers at an 8-byte offset, and with the batch ending exactly at the end of
    the allocation (odd row count), must give the same bits."""
    grid, res, ox, oy = terrain_grid()
    rows, off = sampled_paths(37, seed=91, n_min=1, n_max=300)
    if rows.shape[0] % 2 == 0:                       # odd number of rows: 3 n and n are odd, the last windows are clipped
        rows, off = rows[:-1], np.concatenate([off[:-1], [off[-1] - 1]])
    elev = lookup(grid, res, ox, oy, rows)
    n = rows.shape[0]
    p = shipped_altitude_params()
    dev = torch.device("cuda")
    tool.set_altitude_policy(2)
    d_off = torch.from_numpy(off).to(dev)
    out = {}
    for shift in (0, 1):                             # storage offset in doubles
        buf_r = torch.zeros(3 * n + shift, dtype=torch.float64, device=dev)
        buf_e = torch.zeros(n + shift, dtype=torch.float64, device=dev)
        d_rows, d_elev = buf_r[shift:].view(n, 3), buf_e[shift:]
        d_rows.copy_(torch.from_numpy(rows))
        d_elev.copy_(torch.from_numpy(elev))
        assert d_rows.data_ptr() % 16 == 8 * shift
        d_z1 = torch.empty(n, dtype=torch.float64, device=dev)
        d_solves = torch.zeros(off.shape[0] - 1, dtype=torch.int32, device=dev)
        tool.altitude_optimize_batch_dev(p, d_off, d_rows, d_elev, z_pass1=d_z1, solves=d_solves)
        tool.synchronize()
        out[shift] = (d_rows.cpu().numpy(), d_z1.cpu().numpy(), d_solves.cpu().numpy())
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    host = tool.altitude_optimize_batch(rows, off, p, elev)
    assert np.array_equal(host, out[0][0])
    for b in (0, 5, 36):
        sl = slice(int(off[b]), int(off[b + 1]))
        if sl.stop > sl.start:
            z2 = ao.optimize_segment_altitude_enu(rows[sl], oracle_params(p), elev[sl])
            assert np.abs(out[0][0][sl, 2] - z2).max() <= Z_TOL
