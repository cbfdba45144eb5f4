"""Synthetic terrain and sampled trajectories for the altitude-optimisation tests (the real
raster is a GDAL file outside the tree, so terrain here is an analytic surface rasterised into the cost-map layout of
elevation_cost_map.hpp:49-55: float32 [height, width], row-major, top-left origin, square cells)."""
import numpy as np


def terrain_grid(width=600, height=500, resolution=10.0, origin_x=-1500.0, origin_y=3500.0, seed=0):
    """Rolling hills 900..1500 m; returns (grid float32 [height, width], resolution, origin_x, origin_y)."""
    rng = np.random.default_rng(seed)
    xs = origin_x + (np.arange(width) + 0.5) * resolution
    ys = origin_y - (np.arange(height) + 0.5) * resolution
    X, Y = np.meshgrid(xs, ys)
    z = 1200.0 + 180.0 * np.sin(X / 700.0) * np.cos(Y / 500.0) + 90.0 * np.sin((X + 2 * Y) / 260.0)
    z += rng.normal(0.0, 3.0, z.shape)
    return z.astype(np.float32), resolution, origin_x, origin_y


def sampled_paths(B, seed, n_min=1, n_max=220, spacing=25.0, start_box=((-1200.0, 3800.0), (-1200.0, 3200.0))):
    """Ragged batch of sampled trajectories: a smooth heading random walk with ~`spacing` metres between rows, cruise
    height around the terrain's range (some rows below terrain + clearance, some well above), a few degenerate cases
    (repeated points, rows outside the map).  Returns (rows [n,3], row_offset [B+1])."""
    rng = np.random.default_rng(seed)
    ns = rng.integers(n_min, n_max + 1, B)
    ns[: min(B, 4)] = [1, 2, 3, 4][: min(B, 4)]
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    rows = np.empty((int(off[-1]), 3))
    for b in range(B):
        n = int(ns[b])
        hd = rng.uniform(0, 2 * np.pi) + np.cumsum(rng.normal(0, 0.08, n))
        step = spacing * rng.uniform(0.6, 1.4, n)
        x = rng.uniform(*start_box[0]) + np.cumsum(step * np.cos(hd))
        y = rng.uniform(*start_box[1]) + np.cumsum(step * np.sin(hd))
        z = rng.uniform(1000.0, 1600.0) + np.cumsum(rng.normal(0, 1.5, n))
        if n > 6 and b % 5 == 0:
            x[3], y[3] = x[2], y[2]            # repeated point: the edge is skipped (dist <= 1e-9, cpp:1655)
        rows[off[b]:off[b + 1]] = np.column_stack([x, y, z])
    return rows, off


def lookup(grid, resolution, origin_x, origin_y, rows):
    """Per-row terrain elevation through the oracle's getCostAt; NaN outside the grid."""
    from oracle import alt_oracle as ao

    out = np.full(rows.shape[0], np.nan)
    for i, (x, y, _) in enumerate(rows):
        v = ao.cost_at(grid, resolution, origin_x, origin_y, x, y)
        if v is not None:
            out[i] = v
    return out
