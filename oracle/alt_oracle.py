"""oracle/alt_oracle.py -- TEST INFRASTRUCTURE, not product code.  PARITY UNPINNED.

numpy restatement of the reference's altitude optimiser, the step between the minimum-snap sampler's output and its
conversion to WGS84 in getPlan (/root/reference/uavPathPlanning.cpp:3712-3729 -> runAltitudeOptimization cpp:1535-1573):

    ElevationCostMap::getCostAt           /root/reference/elevation_cost_map.cpp:373-380   (nearest-cell grid lookup)
    UavPathPlanner::optimizeHeights       /root/reference/uavPathPlanning.cpp:1575-1712
    ...::optimizeHeightsGlobalSmooth      /root/reference/uavPathPlanning.cpp:1714-1827
    ...::optimizeSegmentAltitudeENU       /root/reference/uavPathPlanning.cpp:1329-1364   (pass 1, then pass 2 with
                                                                                         lambda_smooth x10, climb x0.5)
    struct AltitudeParams                 /root/reference/uavPathPlanning.hpp:415-421

Each function below follows the reference statement by statement: the Hessian is accumulated triplet by triplet in the
order the reference emplaces them (Eigen's setFromTriplets sums duplicates; the summation order only moves the last
bit of an entry), then solved by a dense Cholesky factorisation.  "Unpinned": the reference solves with Eigen::SimplicialLDLT (cpp:1670, 1796),
Eigen is not installable in this image (no network) and the reference records no outputs for this step, so this port is
not checked against an execution of the reference; it is checked against an independent 50-digit solve of the same
systems (tests/test_alt_oracle.py), which bounds what any backward-stable factorisation -- Eigen's included -- can
return.  The raster itself (GDAL, buildLocalENUCostMap) is out of scope: callers supply the ENU cost grid or the
terrain elevation per point.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np
from scipy.linalg import cho_factor, cho_solve


@dataclass
class AltitudeParams:
    """uavPathPlanning.hpp:415-421 (struct defaults); config.yaml:1-8 ships 1.0 / 1.0 / 0.3 / 2.0 / 10.0."""

    lambda_smooth: float = 1.0
    lambda_follow: float = 0.0
    max_climb_rate: float = 2.0
    uav_R: float = 2.0
    safe_distance: float = 50.0


def shipped_params() -> AltitudeParams:
    """/root/reference/config.yaml:1-8."""
    return AltitudeParams(lambda_smooth=1.0, lambda_follow=1.0, max_climb_rate=0.3, uav_R=2.0, safe_distance=10.0)


def cost_at(grid: np.ndarray, resolution: float, origin_x: float, origin_y: float, x: float, y: float):
    """ElevationCostMap::getCostAt (elevation_cost_map.cpp:373-380): grid is float32 [height, width], row-major,
    top-left origin.  Returns the cell value or None."""
    c = int(np.floor((x - origin_x) / resolution))
    r = int(np.floor((origin_y - y) / resolution))
    if c < 0 or c >= grid.shape[1] or r < 0 or r >= grid.shape[0]:
        return None
    return float(grid[r, c])


def _solve_spd(H, b):
    """H z = b for the SPD Hessian.  The reference factors with Eigen::SimplicialLDLT (cpp:1670, 1796), a Cholesky-type
    method without pivoting for size; LAPACK's dense Cholesky is the closest stand-in available here (an LU with row
    pivoting is visibly noisier on these matrices: their 1e10 / 1e8 penalty rows are harmless to a symmetric
    factorisation but not to row exchanges)."""
    return cho_solve(cho_factor(H, lower=True), b)


def _smooth_and_climb(H, waypoints, p, n):
    """The two blocks both passes share: cpp:1588-1604 / 1741-1757 (smoothing) and cpp:1649-1665 / 1759-1775 (climb)."""
    if n >= 3 and p.lambda_smooth > 0.0:
        s = p.lambda_smooth
        for i in range(1, n - 1):
            H[i - 1, i - 1] += s * 1.0
            H[i - 1, i] += s * -2.0
            H[i - 1, i + 1] += s * 1.0
            H[i, i - 1] += s * -2.0
            H[i, i] += s * 4.0
            H[i, i + 1] += s * -2.0
            H[i + 1, i - 1] += s * 1.0
            H[i + 1, i] += s * -2.0
            H[i + 1, i + 1] += s * 1.0
    if p.max_climb_rate > 0.0:
        for i in range(n - 1):
            dist = float(np.hypot(waypoints[i + 1, 0] - waypoints[i, 0], waypoints[i + 1, 1] - waypoints[i, 1]))
            if dist <= 1e-9:
                continue
            denom = dist * p.max_climb_rate
            if denom <= 1e-12:
                continue
            w = 1.0 / (denom * denom)
            H[i, i] += w
            H[i, i + 1] += -w
            H[i + 1, i] += -w
            H[i + 1, i + 1] += w


def optimize_heights(waypoints, p: AltitudeParams, elev):
    """optimizeHeights, cpp:1575-1712.  ``elev[i]`` is the terrain elevation the map returns at waypoint i (NaN where
    it returns none).  Returns out_z (n,) or None for n == 0."""
    waypoints = np.asarray(waypoints, dtype=np.float64)
    elev = np.asarray(elev, dtype=np.float64)
    n = waypoints.shape[0]
    if n == 0:
        return None
    H = np.zeros((n, n))
    b = np.zeros(n)
    # the reference emplaces smoothing, then follow, then climb, then the regularisation (cpp:1588-1670)
    sm = replace(p, max_climb_rate=0.0)
    _smooth_and_climb(H, waypoints, sm, n)
    for i in range(n):
        if not np.isnan(elev[i]):
            s = p.lambda_follow
            safe_h = elev[i] + p.safe_distance
            target = max(waypoints[i, 2], safe_h)
            H[i, i] += s
            b[i] += s * target
    cl = replace(p, lambda_smooth=0.0)
    _smooth_and_climb(H, waypoints, cl, n)
    for i in range(n):
        H[i, i] += 1e-8
    z = _solve_spd(H, b)
    out = z.copy()
    for i in range(n):                                   # cpp:1684-1709
        if not np.isnan(elev[i]):
            min_h = elev[i] + p.safe_distance
            if out[i] < min_h:
                out[i] = min_h
    return out


def optimize_heights_global_smooth(input_z, waypoints, p: AltitudeParams, return_info: bool = False):
    """optimizeHeightsGlobalSmooth, cpp:1714-1827: at most 10 solves with a growing active set."""
    input_z = np.asarray(input_z, dtype=np.float64)
    waypoints = np.asarray(waypoints, dtype=np.float64)
    n = input_z.shape[0]
    if n == 0 or waypoints.shape[0] != n:
        return None
    current_z = input_z.copy()
    active = np.zeros(n, dtype=bool)
    solves = 0
    for _ in range(10):
        H = np.zeros((n, n))
        b = np.zeros(n)
        _smooth_and_climb(H, waypoints, p, n)
        fix_weight = 1e10
        H[0, 0] += fix_weight
        b[0] += fix_weight * input_z[0]
        H[n - 1, n - 1] += fix_weight
        b[n - 1] += fix_weight * input_z[n - 1]
        constraint_weight = 1e8
        for i in range(1, n - 1):
            if active[i]:
                H[i, i] += constraint_weight
                b[i] += constraint_weight * input_z[i]
        for i in range(n):
            H[i, i] += 1e-8
        z = _solve_spd(H, b)
        solves += 1
        violation = False
        for i in range(n):
            current_z[i] = z[i]
            if current_z[i] < input_z[i] - 1e-3:
                if not active[i]:
                    active[i] = True
                    violation = True
        if not violation:
            break
    np.maximum(current_z, input_z, out=current_z)        # cpp:1817-1819
    return (current_z, solves, active) if return_info else current_z


def optimize_segment_altitude_enu(segment_enu, p: AltitudeParams, elev, return_info: bool = False):
    """optimizeSegmentAltitudeENU, cpp:1329-1364: rows [east, north, up] -> new up values."""
    seg = np.asarray(segment_enu, dtype=np.float64)
    if seg.shape[0] == 0:
        return None
    out_z = optimize_heights(seg, p, elev)
    p2 = replace(p, lambda_smooth=p.lambda_smooth * 10.0, max_climb_rate=p.max_climb_rate * 0.5)
    z2, solves, active = optimize_heights_global_smooth(out_z, seg, p2, return_info=True)
    return (z2, out_z, solves, active) if return_info else z2


# ---- O(n) variant for long trajectories ---------------------------------------------------------------------------
def _band(n, s, w, extra):
    """Lower banded storage (scipy solveh_banded) of the pass Hessian: smoothing (cpp:1588-1604), climb weights w[i] of
    the edges (i, i+1) (cpp:1649-1665), `extra` on the diagonal, 1e-8 regularisation."""
    ab = np.zeros((3, n))
    if n >= 3 and s > 0:
        inner = np.zeros(n + 2)                # inner[i + 1] = 1 if row i is an interior row
        inner[2:n] = 1.0
        ab[0] += s * (inner[2:] + 4.0 * inner[1:-1] + inner[:-2])
        ab[1, :n - 1] += s * -2.0 * (inner[1:n] + inner[2:n + 1])
        ab[2, :n - 2] += s * inner[2:n]
    ab[0, :n - 1] += w
    ab[0, 1:] += w
    ab[1, :n - 1] -= w
    ab[0] += extra + 1e-8
    return ab


def _climb_weights(seg, max_climb_rate):
    dist = np.hypot(np.diff(seg[:, 0]), np.diff(seg[:, 1]))
    denom = dist * max_climb_rate
    ok = (dist > 1e-9) & (denom > 1e-12) & (max_climb_rate > 0.0)
    return np.where(ok, 1.0 / np.where(ok, denom, 1.0) ** 2, 0.0)


def optimize_segment_altitude_enu_banded(segment_enu, p: AltitudeParams, elev, return_info: bool = False):
    """optimize_segment_altitude_enu with vectorised assembly and LAPACK's banded Cholesky (O(n) per solve): the same
    systems and the same active-set loop, entries summed in a different order (last-bit differences).  Checked against
    the statement-by-statement version in tests/test_alt_oracle.py; used where n is in the thousands."""
    from scipy.linalg import solveh_banded

    seg = np.asarray(segment_enu, dtype=np.float64)
    elev = np.asarray(elev, dtype=np.float64)
    n = seg.shape[0]
    if n == 0:
        return None
    has = ~np.isnan(elev)
    w1 = _climb_weights(seg, p.max_climb_rate) if n > 1 else np.zeros(0)
    tgt = np.where(has, np.maximum(seg[:, 2], elev + p.safe_distance), 0.0)
    z1 = solveh_banded(_band(n, p.lambda_smooth, w1, np.where(has, p.lambda_follow, 0.0)),
                       np.where(has, p.lambda_follow * tgt, 0.0), lower=True)
    z1 = np.where(has, np.maximum(z1, elev + p.safe_distance), z1)
    w2 = _climb_weights(seg, p.max_climb_rate * 0.5) if n > 1 else np.zeros(0)
    act = np.zeros(n, dtype=bool)
    cur = z1.copy()
    solves = 0
    for _ in range(10):
        extra = np.zeros(n)
        extra[1:n - 1] = np.where(act[1:n - 1], 1e8, 0.0)
        extra[0] += 1e10
        extra[n - 1] += 1e10
        cur = solveh_banded(_band(n, p.lambda_smooth * 10.0, w2, extra), extra * z1, lower=True)
        solves += 1
        new = (cur < z1 - 1e-3) & ~act
        act |= new
        if not new.any():
            break
    z2 = np.maximum(cur, z1)
    return (z2, z1, solves, act) if return_info else z2
