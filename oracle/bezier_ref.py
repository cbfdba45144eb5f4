"""oracle/bezier_ref.py -- TEST INFRASTRUCTURE, not product code.

ctypes door onto ``oracle/_ref/libbezier_ref*.so`` (built by ``oracle/Makefile`` from ``oracle/bezier_wrapper.cpp``):
the UNMODIFIED reference ``/root/reference/math_util/bezier.cpp`` (class ``math_util::Bezier``) compiled against the
oracle's Eigen shim, and the free helper functions of the single-patrol post-processing cut out of
``/root/reference/uavPathPlanning.cpp:118-206`` at build time (``hasSelfIntersection2D``, ``segmentsIntersect2D``,
``sampleClosedPolygonBoundary``).  Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)
_LIBS: dict = {}


def _path(kind):
    return os.path.join(_HERE, "_ref", "libbezier_ref.so" if kind == "parity" else "libbezier_ref_fast.so")


def available(kind="parity") -> bool:
    return os.path.exists(_path(kind))


def lib(kind="parity"):
    if kind in _LIBS:
        return _LIBS[kind]
    if not available(kind):
        raise FileNotFoundError(f"{_path(kind)} missing: run `make -C oracle` where /root/reference is present")
    L = C.CDLL(_path(kind))
    L.bezier_ref_generate.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_int, _dp]
    L.bezier_ref_generate.restype = C.c_int
    L.bezier_ref_generate_batch.argtypes = [C.c_int, _lp, _dp, C.c_double, C.c_double, C.c_int, _ip, C.c_int, _dp]
    L.bezier_ref_generate_batch.restype = C.c_int
    L.patrol_ref_has_self_intersection.argtypes = [C.c_int, _dp, C.c_int]
    L.patrol_ref_has_self_intersection.restype = C.c_int
    L.patrol_ref_segments_intersect.argtypes = [_dp] * 4
    L.patrol_ref_segments_intersect.restype = C.c_int
    L.patrol_ref_sample_boundary.argtypes = [C.c_int, _dp, C.c_double, C.c_int, _dp]
    L.patrol_ref_sample_boundary.restype = C.c_int
    _LIBS[kind] = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return a.ctypes.data_as(_dp)


def generate(path, sample_distance_override=-1.0, min_radius_arg=0.0, kind="parity"):
    """Bezier::GenerateTrajectoryMatrix as UavPathPlanner::Bezier_3D drives it (cpp:4477-4505): BezierConfig::min_radius =
    300 iff ``min_radius_arg`` > 0, else the default 1.0.  path (n,3) -> rows (S,3); n < 2 -> (0,3)."""
    path = _f64(path)
    n = path.shape[0]
    if n < 2:
        return np.zeros((0, 3))
    cap = 1 << 12
    while True:
        out = np.zeros((cap, 3))
        s = lib(kind).bezier_ref_generate(n, _d(path), float(sample_distance_override), float(min_radius_arg), cap, _d(out))
        if s <= cap:
            return out[:s].copy()
        cap = s


def generate_batch(pt_offset, waypoints, sample_distance_override=-1.0, min_radius_arg=0.0, nthreads=0, cap=0, kind="fast"):
    """B independent calls, OpenMP over trajectories.  Returns (counts [B], threads used, rows [B, cap, 3] or None)."""
    pt_offset = np.ascontiguousarray(pt_offset, dtype=np.int64)
    waypoints = _f64(waypoints)
    B = pt_offset.shape[0] - 1
    counts = np.zeros(B, dtype=np.int32)
    rows = np.zeros((B, cap, 3)) if cap > 0 else None
    used = lib(kind).bezier_ref_generate_batch(B, pt_offset.ctypes.data_as(_lp), _d(waypoints), float(sample_distance_override),
                                               float(min_radius_arg), nthreads, counts.ctypes.data_as(_ip), cap,
                                               _d(rows) if rows is not None else None)
    return counts, used, rows


def has_self_intersection(rows, closed=True) -> bool:
    """hasSelfIntersection2D (uavPathPlanning.cpp:152-177)."""
    rows = _f64(rows)
    return bool(lib().patrol_ref_has_self_intersection(rows.shape[0], _d(rows), int(bool(closed))))


def segments_intersect(a1, a2, b1, b2) -> bool:
    """segmentsIntersect2D (uavPathPlanning.cpp:133-150)."""
    a1, a2, b1, b2 = (_f64(np.asarray(v, dtype=float).reshape(3)) for v in (a1, a2, b1, b2))
    return bool(lib().patrol_ref_segments_intersect(_d(a1), _d(a2), _d(b1), _d(b2)))


def sample_closed_polygon_boundary(polygon, spacing):
    """sampleClosedPolygonBoundary (uavPathPlanning.cpp:179-206)."""
    polygon = _f64(polygon)
    cap = 1 << 12
    while True:
        out = np.zeros((cap, 3))
        m = lib().patrol_ref_sample_boundary(polygon.shape[0], _d(polygon), float(spacing), cap, _d(out))
        if m <= cap:
            return out[:m].copy()
        cap = m
