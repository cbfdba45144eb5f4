"""CPU tests of the altitude-optimisation oracle (oracle/alt_oracle.py, a statement-by-statement numpy restatement of
uavPathPlanning.cpp:1329-1364, 1575-1827).  The reference records no outputs for this step and its solver
(Eigen::SimplicialLDLT) cannot be built here, so the port is checked against an independent 50-digit solve of the same
systems and against the properties the formulation guarantees."""
import mpmath as mp
import numpy as np
import pytest

from alt_helpers import lookup, sampled_paths, terrain_grid
from oracle import alt_oracle as ao


def exact_pass2(z_in, wp, p, active):
    """One pass-2 system (cpp:1737-1797) for a given active set, solved in 50-digit arithmetic."""
    mp.mp.dps = 50
    n = len(z_in)
    H = mp.zeros(n, n)
    b = mp.zeros(n, 1)
    s = mp.mpf(p.lambda_smooth)
    if n >= 3 and p.lambda_smooth > 0:
        for i in range(1, n - 1):
            for a, ca in ((i - 1, 1), (i, -2), (i + 1, 1)):
                for c, cc in ((i - 1, 1), (i, -2), (i + 1, 1)):
                    H[a, c] += s * ca * cc
    for i in range(n - 1):
        dist = float(np.hypot(wp[i + 1, 0] - wp[i, 0], wp[i + 1, 1] - wp[i, 1]))
        if dist <= 1e-9 or dist * p.max_climb_rate <= 1e-12:
            continue
        w = 1 / (mp.mpf(dist) * mp.mpf(p.max_climb_rate)) ** 2
        H[i, i] += w
        H[i + 1, i + 1] += w
        H[i, i + 1] -= w
        H[i + 1, i] -= w
    for k in (0, n - 1):
        H[k, k] += mp.mpf(10) ** 10
        b[k] += mp.mpf(10) ** 10 * mp.mpf(float(z_in[k]))
    for i in range(1, n - 1):
        if active[i]:
            H[i, i] += mp.mpf(10) ** 8
            b[i] += mp.mpf(10) ** 8 * mp.mpf(float(z_in[i]))
    for i in range(n):
        H[i, i] += mp.mpf("1e-8")
    z = mp.lu_solve(H, b)
    return np.array([float(v) for v in z])


def test_cost_at_is_nearest_cell_with_top_left_origin():
    g = np.arange(12, dtype=np.float32).reshape(3, 4)      # height 3, width 4
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 100.0, 50.0) == 0.0          # top-left corner belongs to cell (0, 0)
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 139.99, 20.01) == 11.0       # bottom-right cell
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 115.0, 35.0) == 5.0
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 99.99, 45.0) is None
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 140.0, 45.0) is None
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 105.0, 50.01) is None
    assert ao.cost_at(g, 10.0, 100.0, 50.0, 105.0, 20.0) is None         # y = origin_y - 30: row 3 is outside


def test_port_against_50_digit_solve_and_properties():
    grid, res, ox, oy = terrain_grid()
    rows, off = sampled_paths(10, seed=3, n_min=5, n_max=60)
    p = ao.shipped_params()
    for b in range(10):
        seg = rows[off[b]:off[b + 1]]
        elev = lookup(grid, res, ox, oy, seg)
        z2, z1, solves, active = ao.optimize_segment_altitude_enu(seg, p, elev, return_info=True)
        n = seg.shape[0]
        assert 1 <= solves <= 10
        has = ~np.isnan(elev)
        assert np.all(z1[has] >= elev[has] + p.safe_distance)            # clearance after pass 1 (cpp:1705-1707)
        assert np.all(z2 >= z1)                                           # pass 2 never goes below pass 1 (cpp:1817-1819)
        assert abs(z2[0] - z1[0]) < 1e-6 and abs(z2[-1] - z1[-1]) < 1e-6  # end points held by the 1e10 penalty
        if n >= 3:
            # the port's last solve against exact arithmetic for the same active set
            from dataclasses import replace

            p2 = replace(p, lambda_smooth=p.lambda_smooth * 10, max_climb_rate=p.max_climb_rate * 0.5)
            act_before_last = active.copy()
            z_exact = np.maximum(exact_pass2(z1, seg, p2, act_before_last), z1)
            # `active` after the loop may hold rows added by the last solve only if the loop hit its limit
            if solves < 10:
                assert np.abs(z_exact - z2).max() < 1e-6


@pytest.mark.parametrize("n", [1, 2, 3])
def test_tiny_trajectories(n):
    seg = np.column_stack([np.arange(n) * 30.0, np.zeros(n), np.full(n, 1000.0)])
    elev = np.full(n, 995.0)
    z = ao.optimize_segment_altitude_enu(seg, ao.shipped_params(), elev)
    assert z.shape == (n,) and np.all(z >= 1005.0 - 1e-9) and np.all(np.isfinite(z))
    assert ao.optimize_segment_altitude_enu(np.zeros((0, 3)), ao.shipped_params(), np.zeros(0)) is None


def test_no_terrain_keeps_a_straight_profile():
    seg = np.column_stack([np.arange(40) * 25.0, np.zeros(40), np.linspace(1000.0, 1100.0, 40)])
    z = ao.optimize_segment_altitude_enu(seg, ao.shipped_params(), np.full(40, np.nan))
    assert np.all(np.isfinite(z))


def test_banded_variant_equals_the_statement_by_statement_port():
    grid, res, ox, oy = terrain_grid()
    rows, off = sampled_paths(24, seed=12, n_min=1, n_max=120)
    for p in (ao.shipped_params(), ao.AltitudeParams()):
        for b in range(24):
            seg = rows[off[b]:off[b + 1]]
            elev = lookup(grid, res, ox, oy, seg)
            a = ao.optimize_segment_altitude_enu(seg, p, elev, return_info=True)
            c = ao.optimize_segment_altitude_enu_banded(seg, p, elev, return_info=True)
            assert a[2] == c[2] and np.array_equal(a[3], c[3])           # solves, active set
            assert np.abs(a[1] - c[1]).max() <= 1e-7 and np.abs(a[0] - c[0]).max() <= 1e-7
