"""CPU tests of the multi-GPU host logic: trajectory-index sharding (no data-path collective) and the off-path
exclusive scan of per-rank sample counts, exercised with a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cs_pathplan_b200 import shard_batch, shard_bounds
from cs_pathplan_b200 import workloads
from cs_pathplan_b200.sharding import global_sample_base


@pytest.mark.parametrize("B,world", [(0, 1), (1, 8), (7, 2), (4096, 8), (1 << 20, 8), (10, 3)])
def test_uniform_bounds_cover_exactly_once(B, world):
    bounds = shard_bounds(B, world)
    assert len(bounds) == world and bounds[0][0] == 0 and bounds[-1][1] == B
    for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in bounds]
    assert max(sizes) - min(sizes) <= 1


def test_ragged_bounds_balance_segments_not_counts():
    wp, so = workloads.cfg5(B=4096, seed=1237)
    for world in (2, 4, 8):
        bounds = shard_bounds(4096, world, so)
        assert bounds[0][0] == 0 and bounds[-1][1] == 4096
        seg = [int(so[b1] - so[b0]) for b0, b1 in bounds]
        assert sum(seg) == int(so[-1])
        assert max(seg) - min(seg) <= 2 * 256          # within one longest trajectory of each other
        pieces = [shard_batch(wp, r, world, seg_offset=so) for r in range(world)]
        assert sum(p[0].shape[0] for p in pieces) == wp.shape[0]
        for (w, _, lso, (b0, b1)) in pieces:
            assert lso[0] == 0 and lso[-1] + (b1 - b0) == w.shape[0]
        assert np.array_equal(np.vstack([p[0] for p in pieces]), wp)


def test_uniform_shard_batch_slices():
    wp, ns = workloads.cfg2(B=10, ns=4)
    parts = [shard_batch(wp, r, 3, ns=ns) for r in range(3)]
    assert np.array_equal(np.vstack([p[0] for p in parts]), wp)
    assert [p[3] for p in parts] == shard_bounds(10, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, rows, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wp, ns = workloads.cfg2(B=9, ns=3)
        local, _, _, (b0, b1) = shard_batch(wp, rank, world, ns=ns)
        base, total = global_sample_base(rows[rank])
        t = torch.tensor([b0, b1, local.shape[0], base, total], dtype=torch.int64)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        if rank == 0:
            out.put([g.tolist() for g in gathered])
    finally:
        dist.destroy_process_group()


def test_world2_gloo_shards_and_sample_scan():
    world, rows = 2, [1234, 99]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (b00, b01, n0, base0, tot0), (b10, b11, n1, base1, tot1) = res
    assert (b00, b01, b10, b11) == (0, 4, 4, 9)
    assert n0 == 4 * 4 and n1 == 5 * 4
    assert (base0, base1) == (0, 1234) and tot0 == tot1 == 1333



def _rows_worker(rank, world, port, out):
    """Each rank runs the post-sampler stages (altitude optimisation, ENU -> WGS84) on its shard of a ragged batch of
    sampled trajectories -- here with the CPU oracles standing in for the GPU -- and places its rows in the global
    layout by the off-path exclusive scan; no other communication."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from alt_helpers import sampled_paths
        from cs_pathplan_b200 import shard_rows
        from oracle import alt_oracle as ao
        from oracle import geo

        rows, off = sampled_paths(13, seed=9, n_min=2, n_max=30)
        local, loff, (b0, b1) = shard_rows(rows, off, rank, world)
        z = local.copy()
        for b in range(b1 - b0):
            sl = slice(int(loff[b]), int(loff[b + 1]))
            z[sl, 2] = ao.optimize_segment_altitude_enu(local[sl], ao.shipped_params(), np.full(sl.stop - sl.start, 1000.0))
        lla = geo.enu_to_wgs84_batch(z, geo.README_ORIGIN)
        base, total = global_sample_base(local.shape[0])
        out.put((rank, b0, b1, base, total, lla))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_row_stages_reassemble():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from alt_helpers import sampled_paths
    from oracle import alt_oracle as ao
    from oracle import geo

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rows, off = sampled_paths(13, seed=9, n_min=2, n_max=30)
    full = rows.copy()
    for b in range(13):
        sl = slice(int(off[b]), int(off[b + 1]))
        full[sl, 2] = ao.optimize_segment_altitude_enu(rows[sl], ao.shipped_params(), np.full(sl.stop - sl.start, 1000.0))
    exp = geo.enu_to_wgs84_batch(full, geo.README_ORIGIN)
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == 13          # contiguous trajectory ranges
    assert got[0][3] == 0 and got[1][3] == got[0][5].shape[0] and got[0][4] == got[1][4] == rows.shape[0]
    glued = np.empty_like(exp)
    for _, _, _, base, _, lla in got:
        glued[base: base + lla.shape[0]] = lla
    assert np.array_equal(glued, exp)                                             # sharding never changes a result
    n0, n1 = got[0][5].shape[0], got[1][5].shape[0]
    assert abs(n0 - n1) <= 2 * 30                                                 # balanced on rows (within one longest trajectory per cut)
