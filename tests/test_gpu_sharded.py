"""One batch sharded by trajectory index over several GPUs (BASELINE.json configs[2] and [4]):
every rank r solves `sharding.shard_batch(wp, r, world)` on its own device with its own handle, the shard outputs are
concatenated into the global CSR layout, and the result must be BITWISE equal to the single-GPU result of the whole
batch.  Rank r runs on device r % device_count, so on a one-GPU box the same shard logic runs through several handles of
device 0 and on an N-GPU box through N devices (`gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_sharded.py`).
(pytest -m gpu)"""
import numpy as np
import pytest

from cs_pathplan_b200 import TrajectoryGeneratorTool, sharding, workloads

pytestmark = pytest.mark.gpu


def _device_count():
    import torch

    return torch.cuda.device_count()


def _sharded(wp, cfg, world, ns=None, seg_offset=None):
    ndev = _device_count()
    parts, base, used = [], 0, set()
    for r in range(world):
        wp_l, ns_l, so_l, (b0, b1) = sharding.shard_batch(wp, r, world, ns=ns, seg_offset=seg_offset)
        if b1 == b0:
            continue
        with TrajectoryGeneratorTool(r % ndev) as t:
            res = t.generate_batch(cfg, wp_l, ns=ns_l, seg_offset=so_l)
        used.add(r % ndev)
        parts.append((res, base))
        base += int(res.sample_offset[-1])
    cat = lambda f: np.concatenate([f(p) for p, _ in parts])
    so = np.concatenate([p.sample_offset[:-1] + b for p, b in parts] + [[base]])
    return dict(times=cat(lambda p: p.times), coeff=cat(lambda p: p.coeff), samples=cat(lambda p: p.samples),
                iters=cat(lambda p: p.iters), max_dev=cat(lambda p: p.max_dev), vw_final=cat(lambda p: p.vw_final),
                best_s=cat(lambda p: p.best_s), flags=cat(lambda p: p.flags), stats=cat(lambda p: p.stats),
                sample_offset=so), used


def _assert_equal(whole, parts):
    for k, v in parts.items():
        assert np.array_equal(getattr(whole, k), v), k


@pytest.mark.parametrize("world", [2, 8])
def test_cfg3_shards_equal_single_gpu_bitwise(tool, world):
    wp, ns = workloads.cfg3(B=20000 + world)           # not a multiple of the world size
    cfg = workloads.synthetic_config(4, "shipped")
    whole = tool.generate_batch(cfg, wp, ns=ns)
    parts, used = _sharded(wp, cfg, world, ns=ns)
    _assert_equal(whole, parts)
    assert len(used) == min(world, _device_count())


@pytest.mark.parametrize("world,weights", [(2, "shipped"), (8, "plain")])
def test_cfg5_balanced_shards_equal_single_gpu_bitwise(tool, world, weights):
    wp, so = workloads.cfg5(B=3000)
    cfg = workloads.synthetic_config(4, weights, sample_distance=0.0)
    whole = tool.generate_batch(cfg, wp, seg_offset=so)
    parts, _ = _sharded(wp, cfg, world, seg_offset=so)
    _assert_equal(whole, parts)
    # the split is balanced on segments, not on trajectory counts
    seg = [int(so[b1] - so[b0]) for b0, b1 in sharding.shard_bounds(3000, world, so)]
    assert max(seg) <= 1.1 * (sum(seg) / world) + 256
