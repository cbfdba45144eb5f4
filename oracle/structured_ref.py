"""ctypes binding of oracle/_ref/libstructured_cpu.so: the library's own structured O(ns) algorithm compiled for the host
(oracle/structured_cpu.cpp) -- the "good CPU" baseline of SURVEY.md section 8(d).  Test / bench infrastructure only: nothing
under cs_pathplan_b200/ imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libstructured_cpu.so")
_lib = None


class _Cfg(C.Structure):  # struct msnap_config (include/msnap.h)
    _fields_ = [("order", C.c_int), ("path_weight", C.c_double), ("vel_zero_weight", C.c_double), ("V_avg", C.c_double),
                ("min_time_s", C.c_double), ("sample_distance", C.c_double), ("start_vel", C.c_double * 3),
                ("end_vel", C.c_double * 3), ("start_acc", C.c_double * 3), ("end_acc", C.c_double * 3)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build() -> str:
    subprocess.check_call(["make", "-C", _HERE, "_ref/libstructured_cpu.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.msnap_structured_cpu_generate.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def generate_batch(waypoints, cfg, ns=None, seg_offset=None, sample_distance_override=-1.0, v_avg_override=-1.0,
                   capacity=None, threads=0, samples=True):
    """cfg: any object with the MinimumSnapConfig fields (oracle.msnap_oracle.MinimumSnapConfig or the package's).
    Returns a dict with times, coeff [n_seg, 3, 2*order], max_dev, iters, vw_final, flags, sample_offset, samples."""
    wp = np.ascontiguousarray(waypoints, dtype=np.float64).reshape(-1, 3)
    if seg_offset is not None:
        so = np.ascontiguousarray(seg_offset, dtype=np.int64)
        B, n_seg, nsu = so.shape[0] - 1, int(so[-1]), 0
    else:
        so, nsu = None, int(ns)
        B = wp.shape[0] // (nsu + 1)
        n_seg = B * nsu
    c = _Cfg()
    c.order = int(cfg.order)
    for f in ("path_weight", "vel_zero_weight", "V_avg", "min_time_s", "sample_distance"):
        setattr(c, f, float(getattr(cfg, f)))
    for f in ("start_vel", "end_vel", "start_acc", "end_acc"):
        getattr(c, f)[:] = [float(v) for v in getattr(cfg, f)]
    m = 2 * c.order
    out = {"times": np.empty(n_seg), "coeff": np.empty((n_seg, 3, m)), "max_dev": np.empty(B),
           "iters": np.empty(B, dtype=np.int32), "vw_final": np.empty(B), "flags": np.zeros(B, dtype=np.uint32),
           "sample_offset": np.zeros(B + 1, dtype=np.int64)}
    if capacity is None:  # candidates per segment + first / end point per trajectory
        va = v_avg_override if v_avg_override > 0 else cfg.V_avg
        capacity = int(n_seg * 3 + 2 * B + 16)
        if samples:
            d = np.diff(wp, axis=0)
            seglen = np.sqrt((d * d).sum(1))
            capacity += int((np.maximum(seglen / max(va, 1e-6), cfg.min_time_s) * 10.0 + 2).sum())
    smp = np.empty((capacity if samples else 0, 3))
    rc = lib().msnap_structured_cpu_generate(
        C.byref(c), C.c_double(sample_distance_override), C.c_double(v_avg_override), C.c_longlong(B), C.c_int(nsu), _p(so),
        _p(wp), _p(out["times"]), _p(out["coeff"]), _p(out["max_dev"]), _p(out["iters"]), _p(out["vw_final"]),
        C.c_longlong(capacity if samples else 0), _p(out["sample_offset"]), _p(smp) if samples else None, _p(out["flags"]),
        C.c_int(threads))
    if rc:
        raise RuntimeError(f"msnap_structured_cpu_generate failed ({rc})")
    out["samples"] = smp[: int(out["sample_offset"][-1])] if samples else smp
    return out
