"""The *_dev entry points of include/msnap.h only enqueue work on the handle's stream (no allocation, no host
synchronisation once the workspace has its size): a whole leader chain -- minimum snap + sampler, cost-map lookup, altitude
optimisation, follower formations, ENU -> WGS84 (getPlan, uavPathPlanning.cpp:3684-3729, 3699, 3931-4398) -- is captured
into ONE CUDA graph and replayed; the replay's outputs equal the eagerly launched ones bit for bit."""
import numpy as np
import pytest
import torch

from alt_helpers import terrain_grid
from cs_pathplan_b200 import TrajectoryGeneratorTool, shipped_altitude_params, workloads
from oracle import geo

pytestmark = pytest.mark.gpu


def test_leader_chain_is_captured_into_one_cuda_graph_and_replays_bitwise():
    dev = torch.device("cuda")
    B, ns, n_followers = 96, 16, 4
    wp, _ = workloads.cfg2(B=B, ns=ns, seed=77)
    wp = wp * np.array([6.0, 6.0, 1.0]) + np.array([0.0, 0.0, 1250.0])
    cfg = workloads.synthetic_config(4, "shipped", sample_distance=15.0)
    cfg.V_avg = 40.0
    grid, res, ox, oy = terrain_grid(width=900, height=900, resolution=10.0, origin_x=-4500.0, origin_y=4500.0)
    p = shipped_altitude_params()
    with TrajectoryGeneratorTool(0) as tool:
        cap = tool.sample_bound(cfg, wp, ns=ns)
        d_wp = torch.from_numpy(wp).to(dev)
        d_grid = torch.from_numpy(grid).to(dev)
        d_off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        d_rows = torch.zeros((cap, 3), dtype=torch.float64, device=dev)
        d_elev = torch.zeros(cap, dtype=torch.float64, device=dev)
        d_lla = torch.zeros((cap, 3), dtype=torch.float64, device=dev)
        d_fol = torch.zeros((n_followers * cap, 3), dtype=torch.float64, device=dev)
        d_solves = torch.zeros(B, dtype=torch.int32, device=dev)
        d_flags = torch.zeros(B, dtype=torch.int32, device=dev)
        f_dist, f_rows = tool.formation_parameters()
        outs = (d_off, d_rows, d_elev, d_lla, d_fol, d_solves, d_flags)

        def chain():
            tool.generate_batch_dev(cfg, d_wp, d_off, d_rows, ns=ns, flags=d_flags)
            tool.cost_map_lookup_dev(d_grid, res, ox, oy, d_rows, d_elev, n_rows=d_off[B:])
            tool.altitude_optimize_batch_dev(p, d_off, d_rows, d_elev, solves=d_solves)
            tool.followers_dev(d_rows, d_off, d_fol, 1, f_dist, n_followers, f_rows, "wgs84", geo.README_ORIGIN)
            tool.enu_to_wgs84_dev(geo.README_ORIGIN, d_rows, d_lla, n_rows=d_off[B:])

        s = torch.cuda.Stream()
        tool.set_stream(s.cuda_stream)
        with torch.cuda.stream(s):
            chain()                                     # eager: sizes the workspace, opts the kernels into their shared memory
            chain()
        s.synchronize()
        eager = [o.clone() for o in outs]
        n_rows = int(d_off[B])
        assert n_rows > 0 and not d_flags.any() and int(d_solves.min()) >= 1
        for o in outs:
            o.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):            # capture: nothing runs, any allocation or synchronisation would raise
            chain()
        torch.cuda.synchronize()
        assert int(d_off[B]) == 0                       # (captured, not executed)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        for got, want in zip(outs, eager):
            assert torch.equal(got[:n_rows] if got.shape[0] == cap else got, want[:n_rows] if want.shape[0] == cap else want)
        assert torch.equal(d_fol[: n_followers * n_rows], eager[4][: n_followers * n_rows])
        tool.set_stream(None)
