"""CPU tests of the planner-stage checker oracle/_ref/libplanner_ref.so = the reference's own function definitions for
WGS84 <-> ENU (uavPathPlanning.cpp:894-1108), the altitude optimiser (cpp:1311-1364, 1575-1827) and the follower formations
(cpp:3931-4398), cut out of the reference file at build time and executed (oracle/cut_planner.sh, oracle/planner_wrapper.cpp):
it reproduces the reference's recorded run (readme.md:11-28) and its own golden vectors, and the line-by-line ports
(oracle/geo_port.c, oracle/alt_oracle.py) agree with it -- which is what pins those ports."""
import json
import os

import numpy as np
import pytest

from alt_helpers import lookup, sampled_paths, terrain_grid
from oracle import alt_oracle as ao
from oracle import geo
from oracle import planner_ref as pr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "planner_golden.npz")
needs_ref = pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libplanner_ref.so not built (needs /root/reference)")


def load():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["manifest"]).decode())


@needs_ref
def test_reference_statements_reproduce_the_recorded_run_digit_for_digit():
    enu = pr.wgs84_to_enu_batch(geo.README_WGS84, geo.README_ORIGIN)
    back = pr.enu_to_wgs84_batch(geo.README_ENU, geo.README_ORIGIN)
    for got, want in ((enu, geo.README_ENU), (back, geo.README_WGS84_BACK)):
        assert all(f"{a:.15f}" == f"{b:.15f}" for a, b in zip(got.ravel(), want.ravel()))
    # and the C port is the same function bit for bit, also far from the recorded points
    rng = np.random.default_rng(5)
    pts = np.column_stack([rng.normal(0, 3e4, 2000), rng.normal(0, 3e4, 2000), rng.uniform(-100, 9000, 2000)])
    for origin in (geo.README_ORIGIN, np.array([-70.5, -33.4, 520.0]), np.array([179.99, 78.2, 0.0])):
        lla = pr.enu_to_wgs84_batch(pts, origin)
        assert np.array_equal(lla, geo.enu_to_wgs84_batch(pts, origin))
        assert np.array_equal(pr.wgs84_to_enu_batch(lla, origin), geo.wgs84_to_enu_batch(lla, origin))


@needs_ref
def test_golden_vectors_are_reproduced():
    z, man = load()
    for i, c in enumerate(man["followers"]):
        out = pr.followers(z[f"leader/{c['leader']}"], geo.README_ORIGIN, c["model"], z["starts"][:c["n_followers"]], **c["params"])
        assert np.array_equal(out, z[f"followers/{i}"]), i
    grid, res, ox, oy = terrain_grid()
    for i, c in enumerate(man["altitude"]):
        out, z1, ok = pr.altitude_batch(z[f"alt/{i}/rows"], z[f"alt/{i}/off"], c["params"], grid if c["grid"] else None, res, ox, oy)
        assert np.array_equal(out, z[f"alt/{i}/out"]) and np.array_equal(z1, z[f"alt/{i}/z1"], equal_nan=True) and ok.all()


def test_altitude_port_agrees_with_the_reference_statements():
    """oracle/alt_oracle.py (the line-by-line port the GPU altitude tests use) against the golden outputs of the reference's
    own optimizeSegmentAltitudeENU: <= 1e-6 m (the two differ in the factorisation only: dense Cholesky vs banded LDL')."""
    z, man = load()
    grid, res, ox, oy = terrain_grid()
    worst = 0.0
    for i, c in enumerate(man["altitude"]):
        rows, off = z[f"alt/{i}/rows"], z[f"alt/{i}/off"]
        elev = lookup(grid, res, ox, oy, rows) if c["grid"] else np.full(rows.shape[0], np.nan)
        po = ao.AltitudeParams(*c["params"])
        for b in range(off.shape[0] - 1):
            sl = slice(int(off[b]), int(off[b + 1]))
            z2, z1, _, _ = ao.optimize_segment_altitude_enu(rows[sl], po, elev[sl], return_info=True)
            worst = max(worst, float(np.abs(z2 - z[f"alt/{i}/out"][sl, 2]).max()), float(np.abs(z1 - z[f"alt/{i}/z1"][sl]).max()))
    assert worst <= 1e-6, worst


def test_follower_golden_structure():
    """Properties that need no oracle: models 2-4 start every follower at its own start point with the leader's first up;
    a follower of the V shape sits formation_distance * sqrt(2) * row away from the leader in the ENU plane."""
    z, man = load()
    for i, c in enumerate(man["followers"]):
        out, leader = z[f"followers/{i}"], z[f"leader/{c['leader']}"]
        assert out.shape == (c["n_followers"], leader.shape[0], 3)
        if c["model"] in (2, 3, 4):
            assert np.array_equal(out[:, 0, :2], z["starts"][:c["n_followers"], :2]) and np.all(out[:, 0, 2] == leader[0, 2])
    c = man["followers"][0]
    enu = geo.wgs84_to_enu_batch(z["followers/0"].reshape(-1, 3), geo.README_ORIGIN).reshape(3, -1, 3)
    d = np.hypot(enu[0, :, 0] - z["leader/uav31_0"][:, 0], enu[0, :, 1] - z["leader/uav31_0"][:, 1])
    assert np.abs(d - 50.0 * np.sqrt(2.0)).max() <= 1e-6
