"""ctypes binding of cs_pathplan_b200/libmsnap_b200.so (the C ABI declared in include/msnap.h).

The shared library is the product; this module only declares its signatures.  If the library has not been built
(`cs_pathplan_b200/csrc/build.sh`, or `__graft_entry__.build()`), importing :func:`lib` raises -- there is no
fallback implementation of any kind.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsnap_b200.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_CAPACITY, ERR_ALLOC, ERR_IO = range(7)
FLAG_NONFINITE, FLAG_TRUNCATED = 1, 2


class msnap_config(C.Structure):
    """struct msnap_config (include/msnap.h) == MinimumSnapConfig (minimum_snap.hpp:9-33)."""

    _fields_ = [
        ("order", C.c_int),
        ("path_weight", C.c_double),
        ("vel_zero_weight", C.c_double),
        ("V_avg", C.c_double),
        ("min_time_s", C.c_double),
        ("sample_distance", C.c_double),
        ("start_vel", C.c_double * 3),
        ("end_vel", C.c_double * 3),
        ("start_acc", C.c_double * 3),
        ("end_acc", C.c_double * 3),
    ]


class msnap_altitude_params(C.Structure):
    """struct msnap_altitude_params (include/msnap.h) == AltitudeParams (uavPathPlanning.hpp:415-421)."""

    _fields_ = [("lambda_smooth", C.c_double), ("lambda_follow", C.c_double), ("max_climb_rate", C.c_double),
                ("uav_R", C.c_double), ("safe_distance", C.c_double)]


_vp = C.c_void_p
_ll = C.c_longlong
_i = C.c_int
_d = C.c_double
_cfgp = C.POINTER(msnap_config)

# name -> (restype, argtypes); every symbol include/msnap.h declares
SIGNATURES = {
    "msnap_version": (_i, []),
    "msnap_status_string": (C.c_char_p, [_i]),
    "msnap_last_error": (C.c_char_p, [_vp]),
    "msnap_create": (_i, [_i, C.POINTER(_vp)]),
    "msnap_destroy": (_i, [_vp]),
    "msnap_set_stream": (_i, [_vp, _vp]),
    "msnap_synchronize": (_i, [_vp]),
    "msnap_set_reweight_policy": (_i, [_vp, _i]),
    "msnap_set_host_chunks": (_i, [_vp, _i]),
    "msnap_set_zero_copy": (_i, [_vp, _i]),
    "msnap_launch_count": (_ll, [_vp]),
    "msnap_config_default": (None, [_cfgp]),
    "msnap_config_load_yaml": (_i, [C.c_char_p, _cfgp]),
    "msnap_solve_qp_batch_dev": (_i, [_vp, _i, _d, _d, _ll, _i] + [_vp] * 9),
    "msnap_solve_qp_batch_host": (_i, [_vp, _i, _d, _d, _ll, _i] + [_vp] * 9),
    "msnap_generate_batch_dev": (_i, [_vp, _cfgp, _d, _d, _ll, _i] + [_vp] * 8 + [_ll] + [_vp] * 4),
    "msnap_generate_batch_host": (_i, [_vp, _cfgp, _d, _d, _ll, _i] + [_vp] * 8 + [_ll] + [_vp] * 4),
    "msnap_sample_bound_dev": (_i, [_vp, _cfgp, _d, _ll, _i, _vp, _vp, _vp]),
    "msnap_sample_bound_host": (_i, [_vp, _cfgp, _d, _ll, _i, _vp, _vp, C.POINTER(_ll)]),
    "msnap_generate_one_host": (_i, [_vp, _cfgp, _d, _d, _i, _vp, _ll, _vp, C.POINTER(_ll)]),
    "msnap_wgs84_to_enu_dev": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "msnap_wgs84_to_enu_host": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "msnap_enu_to_wgs84_dev": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "msnap_enu_to_wgs84_host": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "msnap_enu_to_wgs84_counted_dev": (_i, [_vp, _vp, _ll, _vp, _vp, _vp]),
    "msnap_set_sample_frame": (_i, [_vp, _i, _vp]),
    "msnap_set_waypoint_frame": (_i, [_vp, _i, _vp]),
    "msnap_set_geo_exact_trig": (_i, [_vp, _i]),
    "msnap_debug_geo_steps_dev": (_i, [_vp, _vp, _ll, _vp, _vp, _vp]),
    "msnap_altitude_params_default": (None, [C.POINTER(msnap_altitude_params)]),
    "msnap_altitude_optimize_batch_dev": (_i, [_vp, C.POINTER(msnap_altitude_params), _ll, _vp, _ll] + [_vp] * 5),
    "msnap_altitude_optimize_batch_host": (_i, [_vp, C.POINTER(msnap_altitude_params), _ll] + [_vp] * 6),
    "msnap_set_altitude_policy": (_i, [_vp, _i]),
    "msnap_cost_map_lookup_dev": (_i, [_vp, _vp, _i, _i, _d, _d, _d, _ll, _vp, _vp, _vp]),
    "msnap_bezier_generate_batch_dev": (_i, [_vp, _d, _d, _ll, _i, _vp, _vp, _ll, _vp, _vp, _vp]),
    "msnap_bezier_generate_batch_host": (_i, [_vp, _d, _d, _ll, _i, _vp, _vp, _ll, _vp, _vp, _vp]),
    "msnap_patrol_postprocess_dev": (_i, [_vp, _d, _ll, _i, _vp, _vp, _vp, _vp, _ll, _vp, _ll, _vp, _vp, _vp]),
    "msnap_patrol_postprocess_host": (_i, [_vp, _d, _ll, _i, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp]),
    "msnap_formation_distance": (_d, [_d, _d, _d]),
    "msnap_followers_dev": (_i, [_vp, _i, _d, _i, _i, _i, _vp, _vp, _ll, _vp, _vp, _ll, _ll, _vp]),
    "msnap_followers_host": (_i, [_vp, _i, _d, _i, _i, _i, _vp, _vp, _ll, _vp, _vp, _vp]),
    "msnap_profile_begin": (_i, [_vp]),
    "msnap_profile_end": (_i, [_vp, C.c_char_p, _ll]),
    "msnap_debug_phase_clocks": (_i, [_vp, _i, _vp]),
    "msnap_measure_fp64_peak": (_i, [_vp, C.POINTER(_d)]),
}

_LIB = None


def lib():
    """Load libmsnap_b200.so (once).  Raises RuntimeError if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with cs_pathplan_b200/csrc/build.sh (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the minimum-snap path."
        )
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here == the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


class MsnapError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        msg = lib().msnap_status_string(status).decode()
        super().__init__(f"msnap status {status} ({msg})" + (f": {detail}" if detail else ""))
