"""GPU parity of the Bezier generator and the single-patrol post-processing through the C ABI
(msnap_bezier_generate_batch_*, msnap_patrol_postprocess_* in include/msnap.h).

Oracle: the UNMODIFIED reference bezier.cpp and the patrol helper block of uavPathPlanning.cpp:118-206, compiled
(oracle/_ref/libbezier_ref.so) -- golden vectors in tests/golden/bezier_golden.npz, and run live on seeded batches.
Bars: row COUNTS equal (the count of a Bezier segment is the number of accumulated `t += r` steps <= 1, a rounding-sensitive
discrete quantity), rows within 1e-9 m (north-star bar: 1e-6 m; the only difference is CUDA's vs the host's atan2 / sincos /
hypot), trim index / self-intersection verdict / fallback rows identical.  (pytest -m gpu)"""
import json
import os

import numpy as np
import pytest
import torch

from cs_pathplan_b200 import Bezier, BezierConfig, shipped_config, workloads
from cs_pathplan_b200._lib import ERR_CAPACITY, MsnapError
from oracle import bezier_ref as br
from oracle import patrol_port as pp
from oracle import ref

pytestmark = pytest.mark.gpu
ROW_TOL = 1e-9
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bezier_golden.npz")


def golden_cases():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["manifest"]).decode())


def test_bezier_matches_reference_golden(tool):
    z, man = golden_cases()
    worst = 0.0
    for c in man["bezier"]:
        path, want = z[f"{c['name']}/path"], z[f"{c['name']}/rows"]
        got = tool.Bezier_3D(path, c["sample_distance_override"], -1.0, c["min_radius_arg"])
        assert got.shape == want.shape, (c["name"], got.shape, want.shape)
        worst = max(worst, float(np.abs(got - want).max()))
    print(f"\n[bezier] {len(man['bezier'])} golden cases, max |GPU - reference| = {worst:.3e} m")
    assert worst <= ROW_TOL


def test_bezier_class_mirror_and_error_behaviour(tool):
    bz = Bezier(tool)
    assert bz.GenerateTrajectoryMatrix(np.zeros((1, 3)), "").shape == (0, 3)          # bezier.cpp:129-131
    path = workloads.UAV31_0_ENU
    free = bz.GenerateTrajectoryMatrix(path, "", 300.0)
    assert np.abs(free - br.generate(path, 300.0, 0.0)).max() <= ROW_TOL
    bz.SetConfig(BezierConfig(min_radius=300.0))
    tight = bz.GenerateTrajectoryMatrix(path, "", 300.0, 30.0)
    assert np.abs(tight - br.generate(path, 300.0, 1.0)).max() <= ROW_TOL
    # resolution <= 0 means 1.0 (bezier.cpp:133-136)
    small = path[:3] * 0.01
    assert np.array_equal(bz.GenerateTrajectoryMatrix(small, "", -1.0), bz.GenerateTrajectoryMatrix(small, "", 1.0))
    # too little capacity: exact layout, MSNAP_ERR_CAPACITY, flagged
    with pytest.raises(MsnapError) as e:
        tool.bezier_generate_batch(path, ns=6, sample_distance_override=300.0, capacity=50)
    assert e.value.status == ERR_CAPACITY
    off, rows, flags = e.value.partial
    assert off[-1] == free.shape[0] and rows.shape[0] == 50 and flags[0] & 2 and np.abs(rows - free[:50]).max() <= ROW_TOL


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("min_radius_arg", [0.0, 25.0])
def test_bezier_batch_vs_compiled_reference(tool, ragged, min_radius_arg):
    """Seeded batches against the reference run live (OpenMP over trajectories): every trajectory, counts and rows."""
    if ragged:
        wp, so = workloads.cfg5(B=600, seed=77, ns_max=40)
        kw = dict(seg_offset=so)
    else:
        wp, ns = workloads.cfg2(B=1024, seed=78)
        so = np.arange(1025, dtype=np.int64) * ns
        kw = dict(ns=ns)
    B = so.shape[0] - 1
    sd = 0.8
    off, rows, flags = tool.bezier_generate_batch(wp, sample_distance_override=sd, min_radius=300.0 if min_radius_arg > 0 else 1.0, **kw)
    assert not flags.any()
    pt_off = so + np.arange(B + 1)
    counts, _, want = br.generate_batch(pt_off, wp, sd, min_radius_arg, cap=int(np.diff(off).max()) + 4, kind="parity")
    assert np.array_equal(np.diff(off), counts)
    worst = max(float(np.abs(rows[off[b]:off[b + 1]] - want[b, :counts[b]]).max()) for b in range(B))
    print(f"\n[bezier] {B} trajectories / {int(off[-1])} rows vs the compiled reference: counts equal, max row error {worst:.3e} m")
    assert worst <= ROW_TOL
    # batch == singles, device == host, bitwise
    for b in (0, B // 2, B - 1):
        one = tool.bezier_generate_batch(wp[pt_off[b]:pt_off[b + 1]], ns=int(so[b + 1] - so[b]), sample_distance_override=sd,
                                         min_radius=300.0 if min_radius_arg > 0 else 1.0)[1]
        assert np.array_equal(one, rows[off[b]:off[b + 1]])
    dev = torch.device("cuda")
    d_off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    d_rows = torch.empty((int(off[-1]), 3), dtype=torch.float64, device=dev)
    tool.bezier_generate_batch_dev(torch.from_numpy(wp).to(dev), d_off, d_rows, sample_distance_override=sd,
                                   min_radius=300.0 if min_radius_arg > 0 else 1.0,
                                   **({"seg_offset": torch.from_numpy(so).to(dev)} if ragged else {"ns": kw["ns"]}))
    tool.synchronize()
    assert np.array_equal(d_off.cpu().numpy(), off) and np.array_equal(d_rows.cpu().numpy(), rows)


def test_bezier_nonfinite_segment_is_flagged_not_hung(tool):
    wp, ns = workloads.cfg2(B=8, ns=6)
    good = tool.bezier_generate_batch(wp, ns=ns)
    bad = wp.copy()
    bad[3 * 7 + 2, 0] = np.inf
    off, rows, flags = tool.bezier_generate_batch(bad, ns=ns)
    assert flags[3] & 1 and not np.delete(flags, 3).any()
    for b in (0, 2, 4, 7):
        assert np.array_equal(rows[off[b]:off[b + 1]], good[1][good[0][b]:good[0][b + 1]])


def _zones(n_zones, seed):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n_zones):
        n = int(rng.integers(3, 9))
        a = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(20, 400, n)
        z = np.column_stack([r * np.cos(a), r * np.sin(a), rng.uniform(40, 60, n)])
        if t % 2:
            z[:, 1] *= 0.08                                   # thin slivers cross themselves after smoothing
        out.append(z)
    return out


def test_patrol_postprocess_vs_reference_helpers(tool):
    """gen_single_patrol for a ragged batch of closed loops against the port (whose Minisnap_3D, self-intersection test and
    boundary sampling are the reference's own code executed), one distance per run."""
    cfg = shipped_config()
    rcfg = ref.shipped_config()
    for distance, seed in ((5.0, 3), (20.0, 4), (50.0, 5)):
        zones = _zones(40, seed)
        closed = [tool.close_patrol_zone(z) for z in zones]
        wp = np.vstack(closed)
        so = np.concatenate([[0], np.cumsum([c.shape[0] - 1 for c in closed])]).astype(np.int64)
        keep = np.linspace(70.0, 90.0, len(zones))
        res = tool.generate_batch(cfg, wp, seg_offset=so, sample_distance_override=distance, v_avg_override=30.0, stats=False)
        off, rows, flags = tool.patrol_postprocess(wp, res.sample_offset, res.samples, distance, seg_offset=so, keep_up=keep)
        n_fallback = 0
        for b, z in enumerate(zones):
            info = {}
            want = pp.gen_single_patrol(z, distance, rcfg, 30.0, trajectory_enu=np.array([[0.0, 0.0, keep[b]]]), info=info)
            got = rows[off[b]:off[b + 1]]
            assert bool(flags[b] & 4) == info["fallback"], (distance, b)
            assert got.shape == want.shape, (distance, b, got.shape, want.shape, info["best_idx"])
            assert np.abs(got - want).max() <= 1e-6, (distance, b)
            if info["fallback"]:
                assert np.array_equal(got, want)          # boundary sampling is exact arithmetic on the polygon
            n_fallback += info["fallback"]
        assert 0 < n_fallback < len(zones)
        # single-call mirror == batch
        for b in (0, 1, 7):
            one = tool.gen_single_patrol(zones[b], distance, cfg, 30.0, trajectory_enu=np.array([[0.0, 0.0, keep[b]]]))
            assert np.array_equal(one, rows[off[b]:off[b + 1]])
    assert tool.gen_single_patrol(zones[0][:2], 5.0, cfg, 30.0).shape == (0, 3)


def test_patrol_helpers_match_golden_on_device(tool):
    """The intersection verdict and the boundary sampling alone, on the golden polygons (no generator involved): a loop whose
    'samples' are the polygon's own vertices."""
    z, man = golden_cases()
    for c in man["patrol"]:
        poly = z[f"patrol/{c['name']}/polygon"]
        if poly.shape[0] < 3:
            continue
        closed = tool.close_patrol_zone(poly)
        # rows = the polygon followed by P0 again: the trim keeps everything up to the last P0 (distance 0 to the target)
        rows = np.vstack([poly, poly[:1]])
        off, out, flags = tool.patrol_postprocess(closed, np.array([0, rows.shape[0]]), rows, 25.0, ns=closed.shape[0] - 1)
        loop = np.vstack([rows, rows[:1]])
        loop[:, 2] = poly[0, 2]
        want_hit = br.has_self_intersection(loop, True)
        assert bool(flags[0] & 4) == want_hit, c["name"]
        if want_hit:
            want = z[f"patrol/{c['name']}/boundary_25.0"].copy()
            want[:, 2] = poly[0, 2]
            assert np.array_equal(out, want), c["name"]
        else:
            assert np.array_equal(out, loop), c["name"]
