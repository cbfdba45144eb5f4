"""GPU parity of the follower formation trajectories and of the altitude optimiser against the
reference's OWN statements: oracle/_ref/libplanner_ref.so executes the function definitions cut out of
/root/reference/uavPathPlanning.cpp (generateFollowerTrajectories + formation generators, cpp:3931-4398;
optimizeSegmentAltitudeENU, cpp:1329-1364, 1575-1827) -- golden vectors in tests/golden/planner_golden.npz, and live.

Bars: follower rows within 1e-11 deg in lon / lat (1e-6 m) and 1e-6 m in alt of the reference (ENU frame: 1e-9 m);
optimised heights within 1e-6 m.  (pytest -m gpu)"""
import json
import os

import numpy as np
import pytest
import torch

from alt_helpers import sampled_paths, terrain_grid
from cs_pathplan_b200 import AltitudeParams, shipped_altitude_params, workloads
from oracle import geo
from oracle import planner_ref as pr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "planner_golden.npz")


def load():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["manifest"]).decode())


def _kw(params):
    m = {"cfg_max_row": "cfg_uav_formation_max_row", "in_max_row": "in_uav_formation_max_row"}
    return {m.get(k, k): v for k, v in params.items()}


def test_followers_match_reference_golden(tool):
    z, man = load()
    worst_ll = worst_alt = 0.0
    for i, c in enumerate(man["followers"]):
        want = z[f"followers/{i}"]
        got = tool.generateFollowerTrajectories(z[f"leader/{c['leader']}"], geo.README_ORIGIN, c["model"],
                                                z["starts"][:c["n_followers"]], **_kw(c["params"]))
        assert got.shape == want.shape, i
        worst_ll = max(worst_ll, float(np.abs(got[..., :2] - want[..., :2]).max()))
        worst_alt = max(worst_alt, float(np.abs(got[..., 2] - want[..., 2]).max()))
    print(f"\n[followers] {len(man['followers'])} golden cases: max lon/lat error {worst_ll:.3e} deg, max alt error {worst_alt:.3e} m")
    assert worst_ll <= 1e-11 and worst_alt <= 1e-6


@pytest.mark.parametrize("model", [1, 2, 3, 4])
def test_followers_batch_on_sampler_rows_vs_reference(tool, model):
    """A batch of sampled leader trajectories straight from the generator (device resident), 6 followers each, against the
    reference statements run per trajectory; ENU frame against the reference rows converted back."""
    B, ns, F = 48, 16, 6
    wp, _ = workloads.cfg2(B=B, ns=ns)
    wp = wp * np.array([30.0, 30.0, 1.0])
    cfg = workloads.synthetic_config(4, "shipped", sample_distance=25.0)
    cfg.V_avg = 40.0
    res = tool.generate_batch(cfg, wp, ns=ns)
    starts = np.column_stack([109.56 + 0.001 * np.arange(F), 40.867 + 0.0005 * np.arange(F), 10.0 + np.arange(F)])
    d, mr = tool.formation_parameters(cfg_formation_distance=35.0, cfg_uav_formation_max_row=4)
    dev = torch.device("cuda")
    d_rows = torch.from_numpy(res.samples).to(dev)
    d_off = torch.from_numpy(res.sample_offset).to(dev)
    n = res.samples.shape[0]
    d_out = torch.empty((F * n, 3), dtype=torch.float64, device=dev)
    tool.followers_dev(d_rows, d_off, d_out, model, d, F, mr, "wgs84", geo.README_ORIGIN, torch.from_numpy(starts).to(dev))
    tool.synchronize()
    got = d_out.cpu().numpy()
    host = tool.followers_batch(res.samples, res.sample_offset, model, d, F, mr, "wgs84", geo.README_ORIGIN, starts)
    assert np.array_equal(got, host)                                   # device == host path, bitwise
    enu = tool.followers_batch(res.samples, res.sample_offset, model, d, F, mr, "enu")
    for b in range(B):
        r0, r1 = int(res.sample_offset[b]), int(res.sample_offset[b + 1])
        want = pr.followers(res.samples[r0:r1], geo.README_ORIGIN, model, starts, cfg_formation_distance=35.0, cfg_max_row=4)
        blk = got[F * r0: F * r1].reshape(F, r1 - r0, 3)
        assert np.abs(blk[..., :2] - want[..., :2]).max() <= 1e-11 and np.abs(blk[..., 2] - want[..., 2]).max() <= 1e-6, b
        want_enu = geo.wgs84_to_enu_batch(want.reshape(-1, 3), geo.README_ORIGIN).reshape(F, -1, 3)
        t0 = 1 if model >= 2 else 0                                   # (row 0 of models 2-4 is the start point in WGS84 only)
        assert np.abs(enu[F * r0: F * r1].reshape(F, -1, 3)[:, t0:] - want_enu[:, t0:]).max() <= 2e-7, b
        one = tool.generateFollowerTrajectories(res.samples[r0:r1], geo.README_ORIGIN, model, starts, cfg_formation_distance=35.0,
                                                cfg_uav_formation_max_row=4)
        assert np.array_equal(one, blk)                                # batch == singles, bitwise


def test_altitude_matches_reference_statements_golden(tool):
    """optimizeSegmentAltitudeENU as executed from the reference's own text (banded LDL' stand-in for SimplicialLDLT)."""
    z, man = load()
    grid, res, ox, oy = terrain_grid()
    dev = torch.device("cuda")
    worst = 0.0
    for policy in (0, 1, 2):
        tool.set_altitude_policy(policy)
        for i, c in enumerate(man["altitude"]):
            rows, off = z[f"alt/{i}/rows"], z[f"alt/{i}/off"]
            p = AltitudeParams(*c["params"])
            d_rows = torch.from_numpy(rows.copy()).to(dev)
            d_elev = torch.full((rows.shape[0],), float("nan"), dtype=torch.float64, device=dev)
            if c["grid"]:
                tool.cost_map_lookup_dev(torch.from_numpy(grid).to(dev), res, ox, oy, d_rows, d_elev)
            d_z1 = torch.empty(rows.shape[0], dtype=torch.float64, device=dev)
            d_fl = torch.zeros(off.shape[0] - 1, dtype=torch.int32, device=dev)
            tool.altitude_optimize_batch_dev(p, torch.from_numpy(off).to(dev), d_rows, d_elev, z_pass1=d_z1, flags=d_fl)
            tool.synchronize()
            assert not d_fl.cpu().numpy().any()
            worst = max(worst, float(np.abs(d_rows.cpu().numpy() - z[f"alt/{i}/out"]).max()),
                        float(np.abs(d_z1.cpu().numpy() - z[f"alt/{i}/z1"]).max()))
    tool.set_altitude_policy(2)
    print(f"\n[altitude] 4 golden batches x 3 execution forms vs the reference's statements: max height error {worst:.3e} m")
    assert worst <= 1e-6


def test_altitude_full_size_vs_reference_statements(tool):
    """The benchmark-sized sampler output (4 096 trajectories, ~0.8 M rows), shipped parameters, every trajectory."""
    rows, off, elev = workloads.sampled_rows(4096)
    p = shipped_altitude_params()
    # the reference looks the terrain up in a cost map; give it one that reproduces `elev` exactly: one cell per row is not
    # possible, so compare on the analytic terrain rasterised at 5 m and looked up by both sides
    res, ox, oy = 5.0, -100.0, 1100.0
    w, h = int(6700 / res), int(2300 / res)
    xs = ox + (np.arange(w) + 0.5) * res
    grid = np.repeat((1250.0 + 80.0 * np.sin(xs / 400.0))[None, :], h, axis=0).astype(np.float32)
    dev = torch.device("cuda")
    d_rows = torch.from_numpy(rows.copy()).to(dev)
    d_elev = torch.empty(rows.shape[0], dtype=torch.float64, device=dev)
    tool.cost_map_lookup_dev(torch.from_numpy(grid).to(dev), res, ox, oy, d_rows, d_elev)
    tool.altitude_optimize_batch_dev(p, torch.from_numpy(off).to(dev), d_rows, d_elev)
    tool.synchronize()
    want, _, ok = pr.altitude_batch(rows, off, (p.lambda_smooth, p.lambda_follow, p.max_climb_rate, p.uav_R, p.safe_distance),
                                    grid, res, ox, oy, kind="fast")
    assert ok.all()
    err = np.abs(d_rows.cpu().numpy() - want)
    print(f"\n[altitude] 4096 trajectories / {rows.shape[0]} rows vs the reference's statements: max height error {err.max():.3e} m")
    assert err.max() <= 1e-6
