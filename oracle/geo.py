"""oracle/geo.py -- TEST INFRASTRUCTURE, not product code.

ctypes door onto ``oracle/_ref/libgeo_port*.so`` = ``oracle/geo_port.c``, the plain-C restatement of the reference's
WGS84 <-> ECEF <-> ENU transforms (``/root/reference/uavPathPlanning.cpp:894-1108``, constants
``uavPathPlanning.hpp:134-173``).  Pinned on the reference's own recorded run (``readme.md:11-28``), see
``README_WGS84`` / ``README_ENU`` / ``README_WGS84_BACK`` below and ``tests/test_geo_oracle.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this module.
Layouts: WGS84 rows are ``[lon_deg, lat_deg, alt_m]`` (struct WGS84Point), ENU rows ``[east, north, up]`` (ENUPoint).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_LIBS = {}

# ---- the reference's recorded run, /root/reference/readme.md ---------------------------------------------------
# readme.md:11  input "leader_midway_point_wgs84" of case uav31_0
README_WGS84 = np.array([
    [109.56059880227296, 40.86719901015758, 1669.0],
    [109.2995997466117, 40.86719901015758, 1674.0],
    [109.299698988346, 40.84019989401251, 1674.0],
    [109.38269994693026, 40.84019989401251, 1674.0],
    [109.54869918188973, 40.84019989401251, 1674.0],
    [109.54869918188973, 40.86719901015758, 1674.0],
    [109.54869918188973, 40.868098891288774, 1674.0],
])
# the ENU origin is the leader's start with alt forced to 0 (uavPathPlanning.cpp:3643-3644); in this case the leader
# starts at the first midway point (east/north of waypoint 0 print as -5e-11 / -5e-10, up as 1669.0000000008)
README_ORIGIN = np.array([109.56059880227296, 40.86719901015758, 0.0])
# readme.md:14-20  "Enu_waypoint" = wgs84ToENU_Batch(README_WGS84, origin), printed with 15 decimals
README_ENU = np.array([
    [-0.000000000046327, -0.000000000452815, 1669.000000000820137],
    [-22008.910310499257321, 32.799545377501204, 1636.091338242949178],
    [-22009.474804264991690, -2966.281837991115026, 1635.398165184439677],
    [-15007.552345050633448, -2983.825260306681230, 1655.674289593189314],
    [-1003.853909577760191, -2999.001544960936371, 1673.214552272680066],
    [-1003.446472092303907, 0.068179987007966, 1673.921199759593492],
    [-1003.432888336147585, 100.027485618222272, 1673.920415851918733],
])
# readme.md:22-28  "WGS84Point" = enuToWGS84_Batch(README_ENU, origin)
README_WGS84_BACK = np.array([
    [109.560598802272978, 40.867199010157563, 1668.999999999068677],
    [109.299599746611705, 40.867199010157570, 1673.999999998137355],
    [109.299698988345995, 40.840199894012486, 1673.999999998137355],
    [109.382699946930259, 40.840199894012486, 1673.999999998137355],
    [109.548699181889731, 40.840199894012486, 1673.999999998137355],
    [109.548699181889731, 40.867199010157570, 1673.999999998137355],
    [109.548699181889731, 40.868098891288760, 1673.999999998137355],
])


def _lib(fast: bool = False):
    name = "libgeo_port_fast.so" if fast else "libgeo_port.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            # geo_port.c needs nothing but gcc: build it where it is missing (e.g. a box that received the sources only)
            import subprocess

            try:
                subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
            except Exception as e:
                raise RuntimeError(f"{path} is missing and `make -C oracle port` failed: {e}")
        L = C.CDLL(path)
        L.geo_port_wgs84_to_enu_batch.argtypes = [C.c_longlong, _dp, _dp, _dp, C.c_int]
        L.geo_port_wgs84_to_enu_batch.restype = None
        L.geo_port_enu_to_wgs84_batch.argtypes = [C.c_longlong, _dp, _dp, _dp, C.POINTER(C.c_int), C.c_int]
        L.geo_port_enu_to_wgs84_batch.restype = None
        L.geo_port_wgs84_to_ecef.argtypes = [_dp, _dp]
        L.geo_port_wgs84_to_ecef.restype = None
        L.geo_port_ecef_to_wgs84.argtypes = [_dp, _dp]
        L.geo_port_ecef_to_wgs84.restype = C.c_int
        _LIBS[name] = L
    return _LIBS[name]


def _rows(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("expected an [n, 3] array")
    return a


def _ref3(reference) -> np.ndarray:
    r = np.ascontiguousarray(reference, dtype=np.float64).reshape(-1)
    if r.shape[0] != 3:
        raise ValueError("reference must be (lon, lat, alt)")
    return r


def wgs84_to_enu_batch(targets, reference, threads: int = 1, fast: bool = False) -> np.ndarray:
    """UavPathPlanner::wgs84ToENU_Batch (uavPathPlanning.cpp:1085-1095)."""
    t, r = _rows(targets), _ref3(reference)
    out = np.empty_like(t)
    _lib(fast).geo_port_wgs84_to_enu_batch(t.shape[0], t.ctypes.data_as(_dp), r.ctypes.data_as(_dp),
                                           out.ctypes.data_as(_dp), int(threads))
    return out


def enu_to_wgs84_batch(enu, reference, threads: int = 1, fast: bool = False, return_steps: bool = False):
    """UavPathPlanner::enuToWGS84_Batch (uavPathPlanning.cpp:1098-1108).  ``return_steps`` also returns the number of
    fixed-point steps ecefToWGS84 executed per point (cpp:939-949)."""
    e, r = _rows(enu), _ref3(reference)
    out = np.empty_like(e)
    steps = np.zeros(e.shape[0], dtype=np.int32)
    _lib(fast).geo_port_enu_to_wgs84_batch(e.shape[0], e.ctypes.data_as(_dp), r.ctypes.data_as(_dp),
                                           out.ctypes.data_as(_dp), steps.ctypes.data_as(C.POINTER(C.c_int)),
                                           int(threads))
    return (out, steps) if return_steps else out


def wgs84_to_ecef(lla) -> np.ndarray:
    a = np.ascontiguousarray(lla, dtype=np.float64).reshape(3)
    out = np.empty(3)
    _lib().geo_port_wgs84_to_ecef(a.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    return out


def ecef_to_wgs84(ecef) -> np.ndarray:
    a = np.ascontiguousarray(ecef, dtype=np.float64).reshape(3)
    out = np.empty(3)
    _lib().geo_port_ecef_to_wgs84(a.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    return out
