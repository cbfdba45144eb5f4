"""CPU tests of the Bezier / patrol checkers: the compiled reference (oracle/_ref/libbezier_ref.so = unmodified bezier.cpp +
the patrol helper block of uavPathPlanning.cpp:118-206) against the committed golden vectors it generated, and the port
of gen_single_patrol's control flow on top of it."""
import json

import numpy as np
import pytest

from oracle import bezier_ref as br
from oracle import patrol_port as pp
from oracle import ref

GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "bezier_golden.npz")
needs_ref = pytest.mark.skipif(not br.available(), reason="oracle/_ref/libbezier_ref.so not built (needs /root/reference)")


def load():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["manifest"]).decode())


@needs_ref
def test_compiled_reference_reproduces_golden_rows():
    z, man = load()
    assert len(man["bezier"]) >= 20
    for c in man["bezier"]:
        rows = br.generate(z[f"{c['name']}/path"], c["sample_distance_override"], c["min_radius_arg"])
        assert np.array_equal(rows, z[f"{c['name']}/rows"]), c["name"]


@needs_ref
def test_patrol_helpers_reproduce_golden():
    z, man = load()
    for c in man["patrol"]:
        poly = z[f"patrol/{c['name']}/polygon"]
        assert br.has_self_intersection(np.vstack([poly, poly[:1]]), True) == c["self_intersection_closed"]
        assert br.has_self_intersection(poly, False) == c["self_intersection_open"]
        assert np.array_equal(br.sample_closed_polygon_boundary(poly, 25.0), z[f"patrol/{c['name']}/boundary_25.0"])


def test_golden_structure_and_analytic_properties():
    """Properties of the reference's Bezier output that need no oracle: first row = first waypoint, every segment's
    t = 0 point is the waypoint itself, fewer than 2 points give an empty matrix, a d < 0.1 m segment contributes its end
    waypoint only."""
    z, man = load()
    for c in man["bezier"]:
        path, rows = z[f"{c['name']}/path"], z[f"{c['name']}/rows"]
        if c["name"] not in ("short_segment_first", "vertical_only"):   # (their first segment is a d < 0.1 m fallback)
            assert np.array_equal(rows[0], path[0]), c["name"]
        assert rows.shape[1] == 3 and rows.shape[0] >= path.shape[0] - 1
    path, rows = z["vertical_only/path"], z["vertical_only/rows"]
    assert np.array_equal(rows, path[1:])                    # both segments shorter than 0.1 m in the plane


@needs_ref
def test_gen_single_patrol_port_properties():
    cfg = ref.shipped_config()
    sq = np.array([[0, 0, 50.0], [3000, 0, 52], [3000, 2000, 51], [0, 2000, 49]])
    info = {}
    p = pp.gen_single_patrol(sq, 300.0, cfg, 30.0, trajectory_enu=np.array([[1.0, 2.0, 77.0]]), info=info)
    assert not info["fallback"] and p.shape[0] == info["best_idx"] + 2
    assert np.all(p[:, 2] == 77.0) and np.array_equal(p[0], p[-1])                 # levelled and closed
    assert np.array_equal(p[:-1, :2], info["full"][: info["best_idx"] + 1, :2])
    assert pp.gen_single_patrol(sq[:2], 300.0, cfg, 30.0).shape == (0, 3)          # cpp:1834-1837
    # small / thin zones at a fine spacing cross themselves after smoothing: boundary sampling (cpp:1897-1903)
    rng = np.random.default_rng(3)
    seen = 0
    for t in range(24):
        n = int(rng.integers(3, 9))
        a = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(20, 400, n)
        zone = np.column_stack([r * np.cos(a), r * np.sin(a), rng.uniform(40, 60, n)])
        if t % 2:
            zone[:, 1] *= 0.08
        info = {}
        p = pp.gen_single_patrol(zone, 5.0, cfg, 30.0, info=info)
        if info["fallback"]:
            seen += 1
            assert np.array_equal(p[:, :2], br.sample_closed_polygon_boundary(zone, 5.0)[:, :2]) and np.all(p[:, 2] == zone[0, 2])
        else:
            assert not br.has_self_intersection(p, True) and p.shape[0] == info["best_idx"] + 2
    assert 0 < seen < 24
