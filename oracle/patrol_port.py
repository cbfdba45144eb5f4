"""oracle/patrol_port.py -- TEST INFRASTRUCTURE, not product code.

``UavPathPlanner::gen_single_patrol`` (/root/reference/uavPathPlanning.cpp:1829-1906) restated statement by statement.
It is a member function of the planner's translation unit (yaml-cpp / GDAL / out-of-tree headers: not compilable here), so
the control flow below is a port; everything it CALLS is the reference's own code executed: ``Minisnap_3D`` =
``GenerateTrajectoryMatrix`` of the unmodified minimum_snap.cpp (oracle/ref.py) and the helper functions
``hasSelfIntersection2D`` / ``sampleClosedPolygonBoundary`` cut out of the reference file at build time
(oracle/bezier_ref.py).  Pinned by tests/golden/bezier_golden.npz for the helpers; the trim loop is 15 lines.
"""
from __future__ import annotations

import numpy as np

from . import bezier_ref as br
from . import ref


def gen_single_patrol(patrol_zone, distance, cfg: ref.RefConfig, leader_speed, trajectory_enu=None, info=None):
    """Returns the patrol path rows (P,3); ``info`` (a dict) receives best_idx / fallback / full rows."""
    zone = np.asarray(patrol_zone, dtype=np.float64).reshape(-1, 3)
    if zone.shape[0] < 3:                                                    # cpp:1834-1837
        return np.zeros((0, 3))
    te = None if trajectory_enu is None else np.asarray(trajectory_enu, dtype=np.float64).reshape(-1, 3)
    keep_up = te[-1, 2] if te is not None and te.shape[0] else zone[0, 2]    # cpp:1839
    wps = np.vstack([zone, zone[:1]])                                        # cpp:1841-1842
    if wps.shape[0] > 2:
        wps = np.vstack([wps, wps[1:2]])                                     # cpp:1845-1847
    full = ref.generate(wps, cfg, distance, leader_speed)                    # Minisnap_3D, cpp:1849 (4440-4474)
    if full.shape[0] == 0:
        return np.zeros((0, 3))
    if wps.shape[0] > 2:                                                     # cpp:1857-1882
        target = wps[-2]
        best_idx = full.shape[0] - 1
        min_dist = np.finfo(np.float64).max
        search_start = full.shape[0] // 2
        i = full.shape[0]
        while i > search_start:
            i -= 1
            dx, dy, dz = full[i, 0] - target[0], full[i, 1] - target[1], full[i, 2] - target[2]
            d = dx * dx + dy * dy + dz * dz
            if d < min_dist:
                min_dist, best_idx = d, i
        path = full[:best_idx + 1].copy()
    else:
        best_idx = full.shape[0] - 1
        path = full.copy()
    path[:, 2] = keep_up                                                     # cpp:1885-1888
    if path.shape[0]:
        path = np.vstack([path, path[:1]])                                   # cpp:1890-1891
    fallback = br.has_self_intersection(path, True)                          # cpp:1897
    if fallback:
        path = br.sample_closed_polygon_boundary(zone, distance)             # cpp:1899-1902
        path[:, 2] = keep_up
    if info is not None:
        info.update(best_idx=best_idx, fallback=bool(fallback), full=full, closed_waypoints=wps)
    return path
