"""oracle/ref.py -- TEST INFRASTRUCTURE, not product code.

ctypes door onto ``oracle/_ref/libmsnap_ref*.so`` = the UNMODIFIED reference translation unit
``/root/reference/math_util/minimum_snap.cpp`` compiled against ``oracle/shim/Eigen/Dense`` (see
``oracle/Makefile``).  This is "Oracle A" of SURVEY.md section 8(c): same source, same operation order, same
pivoting rule as the reference; only the dense-product summation order differs from real Eigen.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


@dataclass
class RefConfig:
    """Field-for-field mirror of ``MinimumSnapConfig`` (minimum_snap.hpp:9-33); defaults are the struct's."""

    order: int = 3
    path_weight: float = 0.0
    vel_zero_weight: float = 0.0
    V_avg: float = 5.0
    min_time_s: float = 0.1
    sample_distance: float = 1.0
    start_vel: tuple = (0.0, 0.0, 0.0)
    end_vel: tuple = (0.0, 0.0, 0.0)
    start_acc: tuple = (0.0, 0.0, 0.0)
    end_acc: tuple = (0.0, 0.0, 0.0)

    def bc(self) -> np.ndarray:
        return np.ascontiguousarray(
            np.concatenate([self.start_vel, self.end_vel, self.start_acc, self.end_acc]), dtype=np.float64
        )


def shipped_config(**over) -> RefConfig:
    """The values in /root/reference/math_util/minimum_snap_config.yaml:5-27."""
    c = RefConfig(order=2, vel_zero_weight=0.01, path_weight=1e-7, V_avg=200.0, min_time_s=1.0, sample_distance=300.0)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def _host_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return {"avx512f", "avx512vl", "avx512bw", "avx512dq", "avx512cd"} <= fl
    except OSError:
        pass
    return False


def available(kind: str = "parity") -> bool:
    return os.path.exists(_lib_path(kind))


def _lib_path(kind: str) -> str:
    if kind == "parity":
        name = "libmsnap_ref.so"
    elif kind == "fast":
        name = "libmsnap_ref_v4.so" if _host_has_avx512() else "libmsnap_ref_v3.so"
    else:
        raise ValueError(kind)
    return os.path.join(_REF_DIR, name)


_LIBS: dict = {}


def lib(kind: str = "parity"):
    """Load (once) the parity build or the fastest CPU-baseline build this host can run."""
    if kind in _LIBS:
        return _LIBS[kind]
    path = _lib_path(kind)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} missing: run `make -C oracle` where /root/reference is present (see oracle/Makefile)"
        )
    L = C.CDLL(path)
    L.msnap_ref_num_threads.restype = C.c_int
    L.msnap_ref_solve_qp.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp]
    L.msnap_ref_solve_qp.restype = C.c_int
    L.msnap_ref_generate.argtypes = [C.c_int] + [C.c_double] * 5 + [_dp, C.c_double, C.c_double, C.c_int, _dp, C.c_int, _dp]
    L.msnap_ref_generate.restype = C.c_int
    L.msnap_ref_reweighted_solve.argtypes = (
        [C.c_int] + [C.c_double] * 4 + [_dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _ip, _dp]
    )
    L.msnap_ref_reweighted_solve.restype = C.c_int
    L.msnap_ref_generate_batch.argtypes = (
        [C.c_int] + [C.c_double] * 5 + [_dp, C.c_double, C.c_double, C.c_int, _lp, _dp, C.c_int, _ip, C.c_int, _dp]
    )
    L.msnap_ref_generate_batch.restype = C.c_int
    L.msnap_ref_reweighted_solve_batch.argtypes = (
        [C.c_int] + [C.c_double] * 4 + [_dp, C.c_double, C.c_int, _lp, _dp, C.c_int, _dp, _dp, _dp, _ip, _dp]
    )
    L.msnap_ref_reweighted_solve_batch.restype = C.c_int
    _LIBS[kind] = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def solve_qp(order, path, vel, acc, time, path_weight=0.0, vel_zero_weight=0.0, kind="parity"):
    """TrajectoryGeneratorTool::SolveQPClosedForm (ms.cpp:227-649).

    path (n,3), vel/acc (2,3), time (n-1,)  ->  (coeff (ns, 3, 2*order) highest power first, max_dev)."""
    path, vel, acc, time = _f64(path), _f64(vel), _f64(acc), _f64(time)
    n = path.shape[0]
    ns = n - 1
    assert time.shape == (ns,)
    out = np.zeros((ns, 3, 2 * order))
    md = C.c_double(0.0)
    rc = lib(kind).msnap_ref_solve_qp(
        order, n, _d(path), _d(vel), _d(acc), _d(time), path_weight, vel_zero_weight, _d(out), C.byref(md)
    )
    assert rc == 0
    return out, md.value


def generate(path, cfg: RefConfig, sample_distance_override=-1.0, v_avg_override=-1.0, cap=None, kind="parity"):
    """TrajectoryGeneratorTool::GenerateTrajectoryMatrix (ms.cpp:22-206)  ->  samples (S,3)."""
    path = _f64(path)
    n = path.shape[0]
    bc = cfg.bc()
    if cap is None:
        cap = 1 << 16
    while True:
        out = np.zeros((cap, 3))
        s = lib(kind).msnap_ref_generate(
            cfg.order, cfg.path_weight, cfg.vel_zero_weight, cfg.V_avg, cfg.min_time_s, cfg.sample_distance,
            _d(bc), sample_distance_override, v_avg_override, n, _d(path), cap, _d(out),
        )
        if s <= cap:
            return out[:s].copy()
        cap = s


@dataclass
class Reweighted:
    time: np.ndarray
    coeff: np.ndarray
    max_dev: float
    iters: int
    vw_final: float


def reweighted_solve(path, cfg: RefConfig, v_avg_override=-1.0, kind="parity") -> Reweighted:
    """Time allocation (ms.cpp:63-72) + the reweighting loop (ms.cpp:76-90) around the reference's own
    SolveQPClosedForm: the state GenerateTrajectoryMatrix computes internally but does not return."""
    path = _f64(path)
    n = path.shape[0]
    ns = n - 1
    bc = cfg.bc()
    t = np.zeros(ns)
    co = np.zeros((ns, 3, 2 * cfg.order))
    md = C.c_double(0.0)
    it = C.c_int(0)
    vwf = C.c_double(0.0)
    rc = lib(kind).msnap_ref_reweighted_solve(
        cfg.order, cfg.path_weight, cfg.vel_zero_weight, cfg.V_avg, cfg.min_time_s, _d(bc), v_avg_override,
        n, _d(path), _d(t), _d(co), C.byref(md), C.byref(it), C.byref(vwf),
    )
    assert rc == 0
    return Reweighted(t, co, md.value, it.value, vwf.value)


def generate_batch(pt_offset, waypoints, cfg: RefConfig, sample_distance_override=-1.0, v_avg_override=-1.0,
                   nthreads=1, cap=0, kind="fast"):
    """CPU-baseline driver: one GenerateTrajectoryMatrix per trajectory, OpenMP over trajectories.

    Returns (counts[B], threads_used, samples or None)."""
    pt_offset = np.ascontiguousarray(pt_offset, dtype=np.int64)
    waypoints = _f64(waypoints)
    B = pt_offset.shape[0] - 1
    bc = cfg.bc()
    counts = np.zeros(B, dtype=np.int32)
    samples = np.zeros((B, cap, 3)) if cap > 0 else None
    used = lib(kind).msnap_ref_generate_batch(
        cfg.order, cfg.path_weight, cfg.vel_zero_weight, cfg.V_avg, cfg.min_time_s, cfg.sample_distance, _d(bc),
        sample_distance_override, v_avg_override, B, pt_offset.ctypes.data_as(_lp), _d(waypoints), nthreads,
        counts.ctypes.data_as(_ip), cap, _d(samples) if samples is not None else None,
    )
    return counts, used, samples


def reweighted_solve_batch(pt_offset, waypoints, cfg: RefConfig, v_avg_override=-1.0, nthreads=0, kind="parity"):
    """`reweighted_solve` for B trajectories (CSR pt_offset[B+1] into waypoint rows), OpenMP over trajectories.

    Returns (times [sum ns], coeff [sum ns, 3, 2*order], max_dev [B], iters [B], vw_final [B])."""
    pt_offset = np.ascontiguousarray(pt_offset, dtype=np.int64)
    waypoints = _f64(waypoints)
    B = pt_offset.shape[0] - 1
    n_seg = int(pt_offset[-1] - pt_offset[0]) - B
    bc = cfg.bc()
    t = np.zeros(n_seg)
    co = np.zeros((n_seg, 3, 2 * cfg.order))
    md = np.zeros(B)
    it = np.zeros(B, dtype=np.int32)
    vwf = np.zeros(B)
    assert pt_offset[0] == 0
    lib(kind).msnap_ref_reweighted_solve_batch(
        cfg.order, cfg.path_weight, cfg.vel_zero_weight, cfg.V_avg, cfg.min_time_s, _d(bc), v_avg_override, B,
        pt_offset.ctypes.data_as(_lp), _d(waypoints), nthreads, _d(t), _d(co), _d(md), it.ctypes.data_as(_ip), _d(vwf),
    )
    return t, co, md, it, vwf


# readme.md:14-20 -- the only pinned numbers in the reference: ENU waypoints of the uav31_0 leader route
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
