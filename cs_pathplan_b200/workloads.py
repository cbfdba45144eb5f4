"""Synthetic inputs of the BASELINE.json configurations.  Host-side numpy only.

cfg1  uav31_0: the seven ENU leader waypoints printed in the reference's readme.md:14-20
cfg2  4 096 trajectories x 16 segments, order 4        (rng 1234)
cfg3  2^20 trajectories x 8 segments, order 4           (rng 1235)
cfg4  1 024 boustrophedon patrols x 512 segments        (rng 1236)
cfg5  65 536 trajectories, ns log-uniform in [2, 256]   (rng 1237), dense 10 Hz sampling (sample_distance 0)
"""
from __future__ import annotations

import numpy as np

from .api import MinimumSnapConfig, shipped_config

# readme.md:14-20 -- the only numbers the reference pins for this path
UAV31_0_ENU = np.array(
    [
        [-0.000000000046327, -0.000000000452815, 1669.000000000820137],
        [-22008.910310499257321, 32.799545377501204, 1636.091338242949178],
        [-22009.474804264991690, -2966.281837991115026, 1635.398165184439677],
        [-15007.552345050633448, -2983.825260306681230, 1655.674289593189314],
        [-1003.853909577760191, -2999.001544960936371, 1673.214552272680066],
        [-1003.446472092303907, 0.068179987007966, 1673.921199759593492],
        [-1003.432888336147585, 100.027485618222272, 1673.920415851918733],
    ]
)

PLAIN = dict(path_weight=0.0, vel_zero_weight=0.0)
SHIPPED_WEIGHTS = dict(path_weight=1e-7, vel_zero_weight=0.01)  # minimum_snap_config.yaml:7,10


def synthetic_config(order: int = 4, weights: str = "shipped", sample_distance: float = 1.0) -> MinimumSnapConfig:
    """Order-`order` config used by cfg2..cfg5: V_avg 5 m/s, min_time 0.1 s (T about 0.5-7 s, where the reference's
    dense fp64 arithmetic is still meaningful at order 4), `weights` in {"plain", "shipped"}."""
    w = PLAIN if weights == "plain" else SHIPPED_WEIGHTS
    return MinimumSnapConfig(order=order, V_avg=5.0, min_time_s=0.1, sample_distance=sample_distance, **w)


def random_walks(B: int, ns: int, seed: int, sigma: float = 10.0) -> np.ndarray:
    """[B*(ns+1), 3] waypoints: P0 ~ U(-100,100)^3, steps ~ N(0, sigma^2) per axis."""
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-100.0, 100.0, (B, 1, 3))
    steps = rng.normal(0.0, sigma, (B, ns, 3))
    return np.concatenate([p0, p0 + np.cumsum(steps, axis=1)], axis=1).reshape(-1, 3)


def cfg1():
    """(waypoints [7,3], config, sample_distance_override, v_avg_override) as getPlan calls Minisnap_3D."""
    return UAV31_0_ENU.copy(), shipped_config(), 300.0, 30.0


def cfg2(B: int = 4096, ns: int = 16, seed: int = 1234):
    return random_walks(B, ns, seed), ns


def cfg3(B: int = 1 << 20, ns: int = 8, seed: int = 1235):
    return random_walks(B, ns, seed), ns


def cfg4(B: int = 1024, ns: int = 512, seed: int = 1236) -> tuple:
    """Boustrophedon patrol lanes: length 200 m, spacing 20 m, a waypoint every 25 m, jitter N(0, 0.5^2) m,
    z ramp +-0.5 m per lane."""
    rng = np.random.default_rng(seed)
    per_lane = 8  # 200 / 25
    k = np.arange(ns + 1)
    lane, i = k // (per_lane + 1), k % (per_lane + 1)
    x = np.where(lane % 2 == 0, i * 25.0, 200.0 - i * 25.0)
    y = lane * 20.0
    z = 50.0 + 0.5 * lane * np.where(lane % 2 == 0, 1.0, -1.0)
    base = np.stack([x, y, z], axis=1)[None]
    origin = rng.uniform(-500.0, 500.0, (B, 1, 3))
    wp = base + origin + rng.normal(0.0, 0.5, (B, ns + 1, 3))
    return wp.reshape(-1, 3), ns


def cfg5(B: int = 65536, seed: int = 1237, ns_min: int = 2, ns_max: int = 256):
    """Mixed lengths: returns (waypoints [sum ns + B, 3], seg_offset [B+1])."""
    rng = np.random.default_rng(seed)
    ns = np.exp(rng.uniform(np.log(ns_min), np.log(ns_max + 1), B)).astype(np.int64)
    ns = np.clip(ns, ns_min, ns_max)
    seg_offset = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    n_pts = int(seg_offset[-1]) + B
    steps = rng.normal(0.0, 10.0, (n_pts, 3))
    first = seg_offset[:-1] + np.arange(B)
    steps[first] = rng.uniform(-100.0, 100.0, (B, 3))
    # cumulative sum restarted at every trajectory's first point
    cs = np.cumsum(steps, axis=0)
    start_val = cs[first] - steps[first]
    traj_of_pt = np.repeat(np.arange(B), ns + 1)
    wp = cs - start_val[traj_of_pt]
    return wp, seg_offset


def sampled_rows(B: int = 4096, seed: int = 4):
    """A cfg2-sized sampler output for the altitude-optimisation stage: B trajectories of
    150..259 rows at 25 m spacing, cruise height 1300 +- 40 m over an analytic terrain of 1250 +- 80 m.  Returns
    (rows [n,3], row_offset [B+1], elev [n])."""
    rng = np.random.default_rng(seed)
    ns = rng.integers(150, 260, B)
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    n = int(off[-1])
    t = np.arange(n) - np.repeat(off[:-1], ns)
    rows = np.column_stack([t * 25.0, np.repeat(rng.uniform(-1e3, 1e3, B), ns), 1300.0 + 40.0 * np.sin(t / 9.0)])
    elev = 1250.0 + 80.0 * np.sin(rows[:, 0] / 400.0 + np.repeat(rng.uniform(0, 6, B), ns))
    return rows, off, elev


def enu_rows(n: int, seed: int = 11) -> np.ndarray:
    """n ENU rows of a mission area: east/north ~ N(0, 20 km), up ~ U(0, 5 km) (the WGS84 <-> ENU benchmark rows)."""
    rng = np.random.default_rng(seed)
    return np.column_stack([rng.normal(0.0, 2.0e4, n), rng.normal(0.0, 2.0e4, n), rng.uniform(0.0, 5000.0, n)])
