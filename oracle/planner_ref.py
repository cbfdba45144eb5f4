"""oracle/planner_ref.py -- TEST INFRASTRUCTURE, not product code.

ctypes door onto ``oracle/_ref/libplanner_ref*.so`` (``oracle/planner_wrapper.cpp`` + the function definitions that
``oracle/cut_planner.sh`` cuts out of the reference's ``uavPathPlanning.{hpp,cpp}`` / ``elevation_cost_map.cpp`` at build
time): the reference's OWN statements for WGS84 <-> ENU (cpp:894-1108), the altitude optimiser (cpp:1311-1364, 1575-1827)
and the follower formation trajectories (cpp:3931-4398), executed.  Stand-ins: the oracle's Eigen shims (dense; sparse
LDL' in natural order) and a numbers-only json.  Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)
_LIBS: dict = {}


def _path(kind):
    return os.path.join(_HERE, "_ref", "libplanner_ref.so" if kind == "parity" else "libplanner_ref_fast.so")


def available(kind="parity") -> bool:
    return os.path.exists(_path(kind))


def lib(kind="parity"):
    if kind in _LIBS:
        return _LIBS[kind]
    if not available(kind):
        raise FileNotFoundError(f"{_path(kind)} missing: run `make -C oracle` where /root/reference is present")
    L = C.CDLL(_path(kind))
    L.planner_ref_wgs84_to_enu.argtypes = [_dp, C.c_longlong, _dp, _dp]
    L.planner_ref_wgs84_to_enu.restype = None
    L.planner_ref_enu_to_wgs84.argtypes = [_dp, C.c_longlong, _dp, _dp]
    L.planner_ref_enu_to_wgs84.restype = None
    L.planner_ref_altitude_batch.argtypes = [_dp, C.c_int, _lp, _dp, _fp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                             _dp, _ip, C.c_int]
    L.planner_ref_altitude_batch.restype = C.c_int
    L.planner_ref_followers.argtypes = [_dp, C.c_longlong, _dp, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int,
                                        C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, _dp]
    L.planner_ref_followers.restype = C.c_int
    _LIBS[kind] = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return a.ctypes.data_as(_dp)


def wgs84_to_enu_batch(lla, origin, kind="parity"):
    """UavPathPlanner::wgs84ToENU_Batch (cpp:1085-1095)."""
    lla, origin = _f64(lla).reshape(-1, 3), _f64(origin).reshape(3)
    out = np.empty_like(lla)
    lib(kind).planner_ref_wgs84_to_enu(_d(origin), lla.shape[0], _d(lla), _d(out))
    return out


def enu_to_wgs84_batch(enu, origin, kind="parity"):
    """UavPathPlanner::enuToWGS84_Batch (cpp:1098-1108)."""
    enu, origin = _f64(enu).reshape(-1, 3), _f64(origin).reshape(3)
    out = np.empty_like(enu)
    lib(kind).planner_ref_enu_to_wgs84(_d(origin), enu.shape[0], _d(enu), _d(out))
    return out


def altitude_batch(rows, row_offset, params, grid=None, resolution=1.0, origin_x=0.0, origin_y=0.0, nthreads=0, kind="parity"):
    """UavPathPlanner::optimizeSegmentAltitudeENU (cpp:1329-1364) per trajectory of a CSR batch.

    params = (lambda_smooth, lambda_follow, max_climb_rate, uav_R, safe_distance); grid = float32 [height, width] cost map
    (top-left origin at (origin_x, origin_y), square cells) or None.  Returns (rows with the new `up`, z after pass 1, ok[B])."""
    rows = _f64(rows).reshape(-1, 3).copy()
    off = np.ascontiguousarray(row_offset, dtype=np.int64)
    B = off.shape[0] - 1
    p = _f64(params).reshape(5)
    z1 = np.full(rows.shape[0], np.nan)
    ok = np.zeros(B, dtype=np.int32)
    g = None if grid is None else np.ascontiguousarray(grid, dtype=np.float32)
    lib(kind).planner_ref_altitude_batch(_d(p), B, off.ctypes.data_as(_lp), _d(rows), None if g is None else g.ctypes.data_as(_fp),
                                         0 if g is None else g.shape[1], 0 if g is None else g.shape[0], float(resolution),
                                         float(origin_x), float(origin_y), _d(z1), ok.ctypes.data_as(_ip), nthreads)
    return rows, z1, ok


def followers(leader_enu, origin, formation_model, starts_wgs84, cfg_formation_distance=50.0, cfg_position_misalignment=0.0,
              cfg_max_row=8, cfg_uav_R=2.0, in_formation_distance=-1.0, in_position_misalignment=-1.0, in_uav_R=-1.0,
              in_max_row=0, kind="parity"):
    """UavPathPlanner::generateFollowerTrajectories (cpp:3931-4398) for one leader trajectory: returns [F, N, 3] rows
    {lon, lat, alt}.  cfg_* = config.yaml values (hpp:190-199, 185), in_* = the input JSON's overrides (hpp:85-89)."""
    leader = _f64(leader_enu).reshape(-1, 3)
    starts = _f64(starts_wgs84).reshape(-1, 3)
    origin = _f64(origin).reshape(3)
    F, N = starts.shape[0], leader.shape[0]
    out = np.zeros((F, N, 3))
    n = lib(kind).planner_ref_followers(_d(origin), N, _d(leader), int(formation_model), F, _d(starts), float(cfg_formation_distance),
                                        float(cfg_position_misalignment), int(cfg_max_row), float(cfg_uav_R),
                                        float(in_formation_distance), float(in_position_misalignment), float(in_uav_R),
                                        int(in_max_row), _d(out))
    return out[:n]
