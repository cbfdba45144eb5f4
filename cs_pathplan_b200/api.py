"""Host-side mirror of the reference's TrajectoryGeneratorTool interface over the C ABI (include/msnap.h).

Names, argument meaning and error behaviour follow ``/root/reference/math_util/minimum_snap.hpp:9-63``:
``MinimumSnapConfig``, ``TrajectoryGeneratorTool.SolveQPClosedForm`` and
``TrajectoryGeneratorTool.GenerateTrajectoryMatrix``; the batched entry points are the addition this repo makes.
Every compute call goes through ``libmsnap_b200.so`` (CUDA, sm_100a).  Nothing here computes trajectories on the
CPU, and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MsnapError, msnap_altitude_params, msnap_config


@dataclass
class MinimumSnapConfig:
    """minimum_snap.hpp:9-33 -- same field names, same defaults."""

    order: int = 3
    path_weight: float = 0.0
    vel_zero_weight: float = 0.0
    V_avg: float = 5.0
    min_time_s: float = 0.1
    sample_distance: float = 1.0
    start_vel: Sequence[float] = (0.0, 0.0, 0.0)
    end_vel: Sequence[float] = (0.0, 0.0, 0.0)
    start_acc: Sequence[float] = (0.0, 0.0, 0.0)
    end_acc: Sequence[float] = (0.0, 0.0, 0.0)

    def to_c(self) -> msnap_config:
        c = msnap_config()
        c.order = int(self.order)
        c.path_weight = float(self.path_weight)
        c.vel_zero_weight = float(self.vel_zero_weight)
        c.V_avg = float(self.V_avg)
        c.min_time_s = float(self.min_time_s)
        c.sample_distance = float(self.sample_distance)
        for name in ("start_vel", "end_vel", "start_acc", "end_acc"):
            v = [float(x) for x in getattr(self, name)]
            if len(v) != 3:
                raise ValueError(f"{name} must have 3 components")
            setattr(c, name, (C.c_double * 3)(*v))
        return c

    @staticmethod
    def from_c(c: msnap_config) -> "MinimumSnapConfig":
        return MinimumSnapConfig(
            c.order, c.path_weight, c.vel_zero_weight, c.V_avg, c.min_time_s, c.sample_distance,
            tuple(c.start_vel), tuple(c.end_vel), tuple(c.start_acc), tuple(c.end_acc),
        )


@dataclass
class AltitudeParams:
    """struct AltitudeParams, uavPathPlanning.hpp:415-421 -- same field names, same defaults."""

    lambda_smooth: float = 1.0
    lambda_follow: float = 0.0
    max_climb_rate: float = 2.0
    uav_R: float = 2.0
    safe_distance: float = 50.0

    def to_c(self) -> msnap_altitude_params:
        return msnap_altitude_params(float(self.lambda_smooth), float(self.lambda_follow), float(self.max_climb_rate),
                                     float(self.uav_R), float(self.safe_distance))


def shipped_altitude_params() -> AltitudeParams:
    """The altitude_optimization block the reference ships in config.yaml:1-8."""
    return AltitudeParams(lambda_smooth=1.0, lambda_follow=1.0, max_climb_rate=0.3, uav_R=2.0, safe_distance=10.0)


def shipped_config(**over) -> MinimumSnapConfig:
    """The parameters the reference ships in math_util/minimum_snap_config.yaml:5-27."""
    cfg = MinimumSnapConfig(order=2, vel_zero_weight=0.01, path_weight=1e-7, V_avg=200.0, min_time_s=1.0,
                            sample_distance=300.0)
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def load_minimum_snap_config(path: str, base: Optional[MinimumSnapConfig] = None) -> MinimumSnapConfig:
    """Read the min-snap YAML the way UavPathPlanner::loadFromYAML does (uavPathPlanning.cpp:851-879): keys that are
    absent or malformed keep the value they had in ``base`` (default: struct defaults)."""
    L = _lib.lib()
    c = (base or MinimumSnapConfig()).to_c()
    rc = L.msnap_config_load_yaml(path.encode(), C.byref(c))
    if rc != _lib.OK:
        raise MsnapError(rc, path)
    return MinimumSnapConfig.from_c(c)


@dataclass
class BatchResult:
    """Outputs of one batched GenerateTrajectoryMatrix (host arrays)."""

    seg_offset: np.ndarray           # [B+1]
    times: np.ndarray                # [sum ns]
    coeff: np.ndarray                # [sum ns, 3, 2*order]  highest power first
    max_dev: np.ndarray              # [B]
    iters: np.ndarray                # [B] int32
    vw_final: np.ndarray             # [B]
    best_s: np.ndarray               # [sum ns] int32: worst-deviation sample index per segment (0 if path_weight<=0)
    sample_offset: np.ndarray        # [B+1] int64
    samples: np.ndarray              # [rows, 3]
    stats: np.ndarray                # [B, 2]  max climb rate, min turn radius
    flags: np.ndarray                # [B] uint32

    def trajectory(self, b: int) -> np.ndarray:
        return self.samples[self.sample_offset[b]:self.sample_offset[b + 1]]

    def segment_slice(self, b: int) -> slice:
        return slice(int(self.seg_offset[b]), int(self.seg_offset[b + 1]))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a) -> Optional[int]:
    return None if a is None else a.ctypes.data


def _layout(waypoints: np.ndarray, ns: Optional[int], seg_offset):
    """Normalise the (uniform | CSR) batch description.  Returns (B, ns_uniform, seg_offset int64 or None, n_seg)."""
    n_pts = waypoints.shape[0]
    if seg_offset is not None:
        so = np.ascontiguousarray(seg_offset, dtype=np.int64)
        B = so.shape[0] - 1
        if B < 0 or (B >= 0 and so[0] != 0) or np.any(np.diff(so) < 1) or so[-1] + B != n_pts:
            raise ValueError("seg_offset must start at 0, be strictly increasing and cover waypoints (sum ns + B rows)")
        return B, 0, so, int(so[-1])
    if ns is None or ns < 1:
        raise ValueError("give ns (uniform segment count >= 1) or seg_offset")
    if n_pts % (ns + 1) != 0:
        raise ValueError("waypoint rows must be a multiple of ns + 1")
    B = n_pts // (ns + 1)
    return B, int(ns), None, B * int(ns)


class TrajectoryGeneratorTool:
    """GPU-backed stand-in for the reference class of the same name (minimum_snap.hpp:36-63).

    One instance owns one ``msnap_handle`` (device + stream + workspace) and, like the reference object, is not
    re-entrant."""

    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.msnap_create(int(device), C.byref(h))
        if rc != _lib.OK:
            raise MsnapError(rc, f"msnap_create(device={device})")
        self._h = h
        self.device = int(device)

    # ------------------------------------------------------------------ lifetime / plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._L.msnap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, allow=()):
        if rc != _lib.OK and rc not in allow:
            raise MsnapError(rc, self._L.msnap_last_error(self._h).decode())
        return rc

    def set_stream(self, cuda_stream: Optional[int]):
        self._check(self._L.msnap_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._check(self._L.msnap_synchronize(self._h))

    def set_reweight_policy(self, policy: int):
        self._check(self._L.msnap_set_reweight_policy(self._h, int(policy)))

    def set_host_chunks(self, n_chunks: int):
        """0 = automatic pipelining of the host-pointer path, 1 = one chunk (no overlap of copies and kernels)."""
        self._check(self._L.msnap_set_host_chunks(self._h, int(n_chunks)))

    def set_zero_copy(self, enable: bool):
        """Let the kernels store coefficients / samples straight into pinned host buffers (default off: slower than
        the copy engine on B200 / PCIe 5)."""
        self._check(self._L.msnap_set_zero_copy(self._h, int(bool(enable))))

    @property
    def launch_count(self) -> int:
        return int(self._L.msnap_launch_count(self._h))

    def profile_begin(self):
        self._check(self._L.msnap_profile_begin(self._h))

    def profile_end(self) -> dict:
        """Per-kernel {name: {"launches", "total_ms"}} since profile_begin (CUDA events on the launching stream)."""
        import json

        buf = C.create_string_buffer(1 << 16)
        self._check(self._L.msnap_profile_end(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def debug_phase_clocks(self, enable: bool, read: bool = False):
        """Developer instrumentation: arm / read the per-phase clock stamps ([8192,16] int64: rows 0..4095 fused-solve
        CTAs, rows 4096.. sampler tiles)."""
        out = np.zeros((8192, 16), dtype=np.int64) if read else None
        self._check(self._L.msnap_debug_phase_clocks(self._h, int(enable), _ptr(out)))
        return out

    def measure_fp64_peak(self) -> float:
        out = C.c_double(0.0)
        self._check(self._L.msnap_measure_fp64_peak(self._h, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ the reference's two methods
    def SolveQPClosedForm(self, order, Path, Vel, Acc, Time, path_weight=0.0, vel_zero_weight=0.0,
                          return_max_deviation=False):
        """minimum_snap.hpp:45-53.  Path (n,3), Vel/Acc (2,3), Time (n-1,)  ->  PolyCoeff (n-1, 3*2*order)
        (and max_deviation when asked, the reference's optional out-parameter)."""
        Path, Vel, Acc, Time = _f64(Path), _f64(Vel), _f64(Acc), _f64(Time)
        ns = Time.shape[0]
        if Path.ndim != 2 or Path.shape != (ns + 1, 3) or Vel.shape != (2, 3) or Acc.shape != (2, 3) or ns < 1:
            raise ValueError("Path must be (ns+1,3), Vel/Acc (2,3), Time (ns,)")
        coeff, max_dev, _ = self.solve_qp_batch(order, Path, Time, ns=ns, vel=Vel[None], acc=Acc[None],
                                                path_weight=path_weight, vel_zero_weight=vel_zero_weight)
        poly = coeff.reshape(ns, 3 * 2 * order)
        return (poly, float(max_dev[0])) if return_max_deviation else poly

    def GenerateTrajectoryMatrix(self, Path, cfg: MinimumSnapConfig, sample_distance_override=-1.0,
                                 v_avg_override=-1.0) -> np.ndarray:
        """minimum_snap.hpp:60-61.  Path (n,3) -> sampled trajectory (S,3).  An input with fewer than 2 rows or 3
        columns returns an empty (0,0) matrix, as ms.cpp:54-57 does."""
        Path = np.asarray(Path, dtype=np.float64)
        if Path.ndim != 2 or Path.shape[0] < 2 or Path.shape[1] < 3:
            return np.zeros((0, 0))
        wp = _f64(Path[:, :3])
        c = cfg.to_c()
        n = wp.shape[0]
        bound = C.c_longlong(0)
        self._check(self._L.msnap_sample_bound_host(self._h, C.byref(c), v_avg_override, 1, n - 1, None, _ptr(wp),
                                                    C.byref(bound)))
        out = np.empty((max(bound.value, 1), 3))
        cnt = C.c_longlong(0)
        self._check(self._L.msnap_generate_one_host(self._h, C.byref(c), sample_distance_override, v_avg_override, n,
                                                    _ptr(wp), out.shape[0], _ptr(out), C.byref(cnt)))
        return out[:cnt.value].copy()

    # ------------------------------------------------------------------ WGS84 <-> ENU (UavPathPlanner's transforms)
    @staticmethod
    def _geo_args(rows, reference):
        rows = _f64(rows)
        if rows.ndim != 2 or rows.shape[1] != 3:
            raise ValueError("expected [n, 3] rows")
        ref = _f64(reference).reshape(-1)
        if ref.shape[0] != 3:
            raise ValueError("reference must be (lon_deg, lat_deg, alt_m)")
        return rows, ref

    def wgs84ToENU_Batch(self, targets, reference) -> np.ndarray:
        """UavPathPlanner::wgs84ToENU_Batch (uavPathPlanning.cpp:1085-1095): rows [lon, lat, alt] (degrees, metres) ->
        rows [east, north, up] about ``reference`` = (lon, lat, alt)."""
        t, ref = self._geo_args(targets, reference)
        out = np.empty_like(t)
        self._check(self._L.msnap_wgs84_to_enu_host(self._h, _ptr(ref), t.shape[0], _ptr(t), _ptr(out)))
        return out

    def enuToWGS84_Batch(self, targets, reference) -> np.ndarray:
        """UavPathPlanner::enuToWGS84_Batch (uavPathPlanning.cpp:1098-1108): rows [east, north, up] -> [lon, lat, alt]."""
        t, ref = self._geo_args(targets, reference)
        out = np.empty_like(t)
        self._check(self._L.msnap_enu_to_wgs84_host(self._h, _ptr(ref), t.shape[0], _ptr(t), _ptr(out)))
        return out

    def wgs84ToENU(self, target, reference) -> np.ndarray:
        """UavPathPlanner::wgs84ToENU (cpp:1047-1063), one point."""
        return self.wgs84ToENU_Batch(np.asarray(target, dtype=np.float64).reshape(1, 3), reference)[0]

    def enuToWGS84(self, enu, reference) -> np.ndarray:
        """UavPathPlanner::enuToWGS84 (cpp:1066-1083), one point."""
        return self.enuToWGS84_Batch(np.asarray(enu, dtype=np.float64).reshape(1, 3), reference)[0]

    def wgs84_to_enu_dev(self, reference, lla, enu_out):
        """Device rows (CUDA fp64 torch tensors [n,3]); enqueued on the handle's stream; in place allowed."""
        ref = _f64(reference).reshape(3)
        self._check(self._L.msnap_wgs84_to_enu_dev(self._h, _ptr(ref), int(lla.shape[0]), int(lla.data_ptr()),
                                                   int(enu_out.data_ptr())))

    def enu_to_wgs84_dev(self, reference, enu, lla_out, steps_out=None, n_rows=None):
        """n_rows: optional device int64 scalar (e.g. sample_offset[B:]) bounding the rows that are converted."""
        ref = _f64(reference).reshape(3)
        if n_rows is not None:
            self._check(self._L.msnap_enu_to_wgs84_counted_dev(self._h, _ptr(ref), int(enu.shape[0]), int(n_rows.data_ptr()),
                                                               int(enu.data_ptr()), int(lla_out.data_ptr())))
        elif steps_out is None:
            self._check(self._L.msnap_enu_to_wgs84_dev(self._h, _ptr(ref), int(enu.shape[0]), int(enu.data_ptr()),
                                                       int(lla_out.data_ptr())))
        else:
            self._check(self._L.msnap_debug_geo_steps_dev(self._h, _ptr(ref), int(enu.shape[0]), int(enu.data_ptr()),
                                                          int(lla_out.data_ptr()), int(steps_out.data_ptr())))

    def set_geo_exact_trig(self, enable: bool):
        """ENU -> WGS84: run ecefToWGS84's iteration statement by statement with per-step sin/cos/atan2 (True) or on
        direction vectors (False, default; same iteration, ~4x faster, rounding-level differences)."""
        self._check(self._L.msnap_set_geo_exact_trig(self._h, int(bool(enable))))

    def set_sample_frame(self, frame: str = "enu", reference=None):
        """Frame of the sampled rows the generate calls return: "enu" (the reference's GenerateTrajectoryMatrix) or
        "wgs84" = getPlan's enuToWGS84_Batch(Trajectory_ENU, origin_) (cpp:3699) applied on the device."""
        if frame not in ("enu", "wgs84"):
            raise ValueError("frame must be 'enu' or 'wgs84'")
        ref = None if reference is None else _f64(reference).reshape(3)
        self._check(self._L.msnap_set_sample_frame(self._h, 1 if frame == "wgs84" else 0, _ptr(ref)))

    # ------------------------------------------------------------------ altitude optimisation (cpp:1329-1364, 1575-1827)
    def altitude_optimize_batch(self, rows, row_offset, params: AltitudeParams, elev=None, return_info: bool = False):
        """B independent optimizeSegmentAltitudeENU calls: rows [n,3] (east, north, up) with CSR row_offset [B+1], elev [n]
        (terrain elevation per row, NaN = none).  Returns the rows with the optimised ``up`` column (a copy), and with
        ``return_info`` also (z after pass 1 [n], solves of pass 2 [B], flags [B])."""
        rows = _f64(rows).copy()
        off = np.ascontiguousarray(row_offset, dtype=np.int64)
        B = off.shape[0] - 1
        if rows.ndim != 2 or rows.shape[1] != 3 or B < 0 or (B >= 0 and (off[0] != 0 or off[-1] != rows.shape[0])):
            raise ValueError("rows must be [n,3] and row_offset [B+1] with row_offset[0] = 0, row_offset[B] = n")
        elev = None if elev is None else _f64(elev)
        if elev is not None and elev.shape != (rows.shape[0],):
            raise ValueError("elev must have one entry per row")
        z1 = np.empty(rows.shape[0])
        solves = np.zeros(max(B, 0), dtype=np.int32)
        flags = np.zeros(max(B, 0), dtype=np.uint32)
        c = params.to_c()
        self._check(self._L.msnap_altitude_optimize_batch_host(self._h, C.byref(c), B, _ptr(off), _ptr(rows), _ptr(elev),
                                                               _ptr(z1), _ptr(solves), _ptr(flags)))
        return (rows, z1, solves, flags) if return_info else rows

    def optimizeSegmentAltitudeENU(self, segment_enu, params: AltitudeParams, elev=None) -> np.ndarray:
        """UavPathPlanner::optimizeSegmentAltitudeENU (cpp:1329-1364) for one trajectory: returns the rows with the new
        ``up`` values (the reference updates segment_enu in place)."""
        seg = _f64(segment_enu)
        return self.altitude_optimize_batch(seg, np.array([0, seg.shape[0]], dtype=np.int64), params, elev)

    def altitude_optimize_batch_dev(self, params: AltitudeParams, row_offset, rows, elev=None, z_pass1=None, solves=None,
                                    flags=None):
        """Device rows (CUDA torch tensors: row_offset int64 [B+1], rows fp64 [cap,3], elev fp64 [cap]); enqueued on the
        handle's stream, ``up`` column updated in place."""
        c = params.to_c()

        def dp(t):
            return None if t is None else int(t.data_ptr())

        self._check(self._L.msnap_altitude_optimize_batch_dev(
            self._h, C.byref(c), int(row_offset.numel()) - 1, dp(row_offset), int(rows.shape[0]), dp(rows), dp(elev),
            dp(z_pass1), dp(solves), dp(flags)))

    def set_altitude_policy(self, policy: int):
        """0 = lane pairs (two-sided elimination), 1 = one lane per trajectory, 2 = partitioned over 8 / 32 lanes, the
        whole stage in one launch (default)."""
        self._check(self._L.msnap_set_altitude_policy(self._h, int(policy)))

    def cost_map_lookup_dev(self, grid, resolution, origin_x, origin_y, rows, elev_out, n_rows=None):
        """ElevationCostMap::getCostAt (elevation_cost_map.cpp:373-380) for device rows: grid = CUDA float32 tensor
        [height, width] (row-major, top-left origin); n_rows = optional device int64 scalar bounding the rows."""
        self._check(self._L.msnap_cost_map_lookup_dev(
            self._h, int(grid.data_ptr()), int(grid.shape[1]), int(grid.shape[0]), float(resolution), float(origin_x),
            float(origin_y), int(rows.shape[0]), None if n_rows is None else int(n_rows.data_ptr()), int(rows.data_ptr()),
            int(elev_out.data_ptr())))

    def set_waypoint_frame(self, frame: str = "enu", reference=None):
        """Frame of the waypoints the generate / sample_bound calls take: "enu" or "wgs84" rows [lon, lat, alt], converted
        on the device first = prepareWaypoints' wgs84ToENU_Batch(wgs84_points, origin_) (cpp:2640)."""
        if frame not in ("enu", "wgs84"):
            raise ValueError("frame must be 'enu' or 'wgs84'")
        ref = None if reference is None else _f64(reference).reshape(3)
        self._check(self._L.msnap_set_waypoint_frame(self._h, 1 if frame == "wgs84" else 0, _ptr(ref)))

    # ------------------------------------------------------------------ Bezier generator (bezier.cpp:127-189)
    def bezier_generate_batch(self, waypoints, ns=None, seg_offset=None, sample_distance_override=-1.0, min_radius=1.0,
                              capacity: Optional[int] = None):
        """B independent math_util::Bezier::GenerateTrajectoryMatrix calls; host arrays in and out.  ``min_radius`` is
        BezierConfig::min_radius (1.0 = unconstrained).  Returns (sample_offset [B+1], samples [rows,3], flags [B]).
        Without ``capacity`` a sizing call (exact row layout, nothing written) precedes the real one."""
        wp = _f64(waypoints)
        B, nsu, so, _ = _layout(wp, ns, seg_offset)
        off = np.zeros(B + 1, dtype=np.int64)
        flags = np.zeros(B, dtype=np.uint32)

        def call(cap, rows):
            return self._L.msnap_bezier_generate_batch_host(self._h, float(sample_distance_override), float(min_radius), B, nsu,
                                                            _ptr(so), _ptr(wp), int(cap), _ptr(off), _ptr(rows), _ptr(flags))

        if capacity is None:
            self._check(call(0, None), allow=(_lib.ERR_CAPACITY,))
            capacity = int(off[B])
        rows = np.empty((max(int(capacity), 1), 3))
        rc = call(capacity, rows)
        n = int(min(off[B], capacity))
        if rc != _lib.OK:
            err = MsnapError(rc, self._L.msnap_last_error(self._h).decode())
            err.partial = (off, rows[:n], flags)
            raise err
        return off, rows[:n], flags

    def bezier_generate_batch_dev(self, waypoints, sample_offset, samples, ns=None, seg_offset=None,
                                  sample_distance_override=-1.0, min_radius=1.0, flags=None):
        """Device tensors (CUDA torch): enqueued on the handle's stream.  samples may be None (sizing call)."""
        def dp(t):
            return None if t is None else int(t.data_ptr())

        B = int(sample_offset.numel()) - 1
        self._check(self._L.msnap_bezier_generate_batch_dev(
            self._h, float(sample_distance_override), float(min_radius), B, int(ns or 0), dp(seg_offset), dp(waypoints),
            0 if samples is None else int(samples.shape[0]), dp(sample_offset), dp(samples), dp(flags)))

    def Bezier_3D(self, Enu_waypoint, distance, V_avg_override=-1.0, min_radius=0.0) -> np.ndarray:
        """UavPathPlanner::Bezier_3D (uavPathPlanning.cpp:4477-4505): fewer than 2 waypoints -> empty; BezierConfig::
        min_radius = 300 whenever ``min_radius`` > 0 (the planner's own hard-coded value, cpp:4491-4494)."""
        wp = np.asarray(Enu_waypoint, dtype=np.float64).reshape(-1, 3)
        if wp.shape[0] < 2:
            return np.zeros((0, 3))
        _, rows, _ = self.bezier_generate_batch(wp, ns=wp.shape[0] - 1, sample_distance_override=distance,
                                                min_radius=300.0 if min_radius > 0 else 1.0)
        return rows

    # ------------------------------------------------------------------ single-loop patrols (cpp:1829-1906)
    @staticmethod
    def close_patrol_zone(patrol_zone) -> np.ndarray:
        """The waypoint list gen_single_patrol hands to Minisnap_3D: P0..Pn-1, P0, P1 (cpp:1841-1847)."""
        z = np.asarray(patrol_zone, dtype=np.float64).reshape(-1, 3)
        return np.vstack([z, z[:1], z[1:2]])

    def patrol_postprocess(self, waypoints, sample_offset, samples, distance, ns=None, seg_offset=None, keep_up=None,
                           capacity: Optional[int] = None):
        """gen_single_patrol's post-processing of Minisnap_3D's rows for B closed loops (cpp:1857-1903); host arrays.
        ``waypoints`` are the CLOSED lists (close_patrol_zone).  Returns (out_offset [B+1], out_rows, flags [B])."""
        wp = _f64(waypoints)
        B, nsu, so, _ = _layout(wp, ns, seg_offset)
        soff = np.ascontiguousarray(sample_offset, dtype=np.int64)
        rows = _f64(samples).reshape(-1, 3)
        if soff.shape != (B + 1,) or soff[-1] > rows.shape[0]:
            raise ValueError("sample_offset must have B + 1 entries covering `samples`")
        ku = None if keep_up is None else _f64(keep_up).reshape(B)
        off = np.zeros(B + 1, dtype=np.int64)
        flags = np.zeros(B, dtype=np.uint32)

        def call(cap, out):
            return self._L.msnap_patrol_postprocess_host(self._h, float(distance), B, nsu, _ptr(so), _ptr(wp), _ptr(soff),
                                                         _ptr(rows), _ptr(ku), int(cap), _ptr(off), _ptr(out), _ptr(flags))

        if capacity is None:
            self._check(call(0, None), allow=(_lib.ERR_CAPACITY,))
            capacity = int(off[B])
        out = np.empty((max(int(capacity), 1), 3))
        self._check(call(capacity, out))
        return off, out[:int(off[B])], flags

    def patrol_postprocess_dev(self, waypoints, sample_offset, samples, distance, out_offset, out_rows, ns=None,
                               seg_offset=None, keep_up=None, flags=None):
        """Device tensors; enqueued on the handle's stream.  out_rows may be None (sizing call: out_offset only)."""
        def dp(t):
            return None if t is None else int(t.data_ptr())

        B = int(sample_offset.numel()) - 1
        self._check(self._L.msnap_patrol_postprocess_dev(
            self._h, float(distance), B, int(ns or 0), dp(seg_offset), dp(waypoints), dp(sample_offset), dp(samples),
            int(samples.shape[0]), dp(keep_up), 0 if out_rows is None else int(out_rows.shape[0]), dp(out_offset), dp(out_rows),
            dp(flags)))

    def gen_single_patrol(self, patrol_zone, distance, cfg: MinimumSnapConfig, leader_speed: float, trajectory_enu=None):
        """UavPathPlanner::gen_single_patrol (cpp:1829-1906) for one polygon: close it, Minisnap_3D(closed, distance,
        leader_speed), trim / level / close / self-intersection fallback.  ``trajectory_enu``: the rows flown before the
        patrol (its last `up` is kept, cpp:1839).  Fewer than 3 vertices or an empty generator result -> (0,3)."""
        zone = np.asarray(patrol_zone, dtype=np.float64).reshape(-1, 3)
        if zone.shape[0] < 3:
            return np.zeros((0, 3))
        closed = self.close_patrol_zone(zone)
        res = self.generate_batch(cfg, closed, ns=closed.shape[0] - 1, sample_distance_override=distance,
                                  v_avg_override=leader_speed, stats=False)
        te = None if trajectory_enu is None else np.asarray(trajectory_enu, dtype=np.float64).reshape(-1, 3)
        keep = None if te is None or te.shape[0] == 0 else np.array([te[-1, 2]])
        _, rows, _ = self.patrol_postprocess(closed, res.sample_offset, res.samples, distance, ns=closed.shape[0] - 1, keep_up=keep)
        return rows

    # ------------------------------------------------------------------ follower formations (cpp:3931-4398)
    @staticmethod
    def formation_parameters(cfg_formation_distance=50.0, cfg_position_misalignment=0.0, cfg_uav_formation_max_row=8,
                             cfg_uav_R=2.0, in_formation_distance=-1.0, in_position_misalignment=-1.0, in_uav_R=-1.0,
                             in_uav_formation_max_row=0):
        """(formation_distance, uav_formation_max_row) as generateFollowerTrajectories derives them: config.yaml values
        (hpp:185, 190, 198-199) overridden by the input JSON's (hpp:85-89; cpp:4036-4042), max_row >= 1, and the lower bound
        (2 * position_misalignment + uav_R) * 1.41421 on the distance (cpp:4044-4051)."""
        d, pm, mr, r = cfg_formation_distance, cfg_position_misalignment, cfg_uav_formation_max_row, cfg_uav_R
        if in_formation_distance > 0.0:
            d = in_formation_distance
        if in_position_misalignment >= 0.0:
            pm = in_position_misalignment
        if in_uav_R > 0.0:
            r = in_uav_R
        if in_uav_formation_max_row > 0:
            mr = in_uav_formation_max_row
        return float(_lib.lib().msnap_formation_distance(float(d), float(pm), float(r))), max(int(mr), 1)

    def followers_batch(self, leader_rows, row_offset, formation_model, formation_distance, n_followers,
                        uav_formation_max_row=8, frame="wgs84", reference=None, starts_wgs84=None) -> np.ndarray:
        """Follower trajectories of B leader trajectories (rows [n,3] ENU with CSR row_offset [B+1]); host arrays.
        Returns rows [n_followers * n, 3]: trajectory b's block starts at n_followers * row_offset[b], follower-major.
        ``formation_distance`` is the value AFTER the lower bound (formation_parameters)."""
        rows = _f64(leader_rows).reshape(-1, 3)
        off = np.ascontiguousarray(row_offset, dtype=np.int64)
        B = off.shape[0] - 1
        if B < 0 or off[0] != 0 or off[-1] != rows.shape[0]:
            raise ValueError("row_offset must start at 0 and end at the number of rows")
        if frame not in ("enu", "wgs84"):
            raise ValueError("frame must be 'enu' or 'wgs84'")
        ref = None if reference is None else _f64(reference).reshape(3)
        st = None if starts_wgs84 is None else _f64(starts_wgs84).reshape(int(n_followers), 3)
        out = np.empty((int(n_followers) * rows.shape[0], 3))
        self._check(self._L.msnap_followers_host(self._h, int(formation_model), float(formation_distance),
                                                 int(uav_formation_max_row), int(n_followers), 1 if frame == "wgs84" else 0,
                                                 _ptr(ref), _ptr(st), B, _ptr(off), _ptr(rows), _ptr(out)))
        return out

    def followers_dev(self, leader_rows, row_offset, out_rows, formation_model, formation_distance, n_followers,
                      uav_formation_max_row=8, frame="wgs84", reference=None, starts_wgs84=None):
        """Device tensors (leader_rows fp64 [cap,3], row_offset int64 [B+1], out_rows fp64 [>= n_followers * rows, 3],
        starts_wgs84 fp64 [n_followers,3] or None); enqueued on the handle's stream."""
        ref = None if reference is None else _f64(reference).reshape(3)
        self._check(self._L.msnap_followers_dev(
            self._h, int(formation_model), float(formation_distance), int(uav_formation_max_row), int(n_followers),
            1 if frame == "wgs84" else 0, _ptr(ref), None if starts_wgs84 is None else int(starts_wgs84.data_ptr()),
            int(row_offset.numel()) - 1, int(row_offset.data_ptr()), int(leader_rows.data_ptr()), int(leader_rows.shape[0]),
            int(out_rows.shape[0]), int(out_rows.data_ptr())))

    def generateFollowerTrajectories(self, Trajectory_ENU, origin, formation_model, uav_start_point_wgs84, **params):
        """UavPathPlanner::generateFollowerTrajectories (cpp:3931-4074) for one leader trajectory: returns [F, N, 3] rows
        {lon, lat, alt}, one block per follower (the reference's JSON rows without the uav ids).  ``params`` are the
        keyword arguments of ``formation_parameters``."""
        traj = _f64(Trajectory_ENU).reshape(-1, 3)
        starts = _f64(uav_start_point_wgs84).reshape(-1, 3)
        d, mr = self.formation_parameters(**params)
        F, N = starts.shape[0], traj.shape[0]
        if F == 0 or N == 0:
            return np.zeros((F, N, 3))
        out = self.followers_batch(traj, np.array([0, N], dtype=np.int64), formation_model, d, F, mr, "wgs84", origin, starts)
        return out.reshape(F, N, 3)

    # ------------------------------------------------------------------ batched, host buffers
    def solve_qp_batch(self, order, waypoints, times, ns=None, seg_offset=None, vel=None, acc=None,
                       path_weight=0.0, vel_zero_weight=0.0):
        """B closed-form solves (no time allocation, no reweighting).  Returns (coeff [sum ns,3,2o], max_dev [B],
        flags [B])."""
        wp, times = _f64(waypoints), _f64(times)
        B, nsu, so, n_seg = _layout(wp, ns, seg_offset)
        if times.shape != (n_seg,):
            raise ValueError("times must have one entry per segment")
        vel = None if vel is None else _f64(vel).reshape(B, 2, 3)
        acc = None if acc is None else _f64(acc).reshape(B, 2, 3)
        coeff = np.empty((n_seg, 3, 2 * int(order)))
        max_dev = np.empty(B)
        flags = np.zeros(B, dtype=np.uint32)
        best_s = np.zeros(n_seg, dtype=np.int32)
        self._check(self._L.msnap_solve_qp_batch_host(
            self._h, int(order), float(path_weight), float(vel_zero_weight), B, nsu, _ptr(so), _ptr(wp), _ptr(vel),
            _ptr(acc), _ptr(times), _ptr(coeff), _ptr(max_dev), _ptr(best_s), _ptr(flags)))
        self.last_best_s = best_s
        return coeff, max_dev, flags

    def sample_bound(self, cfg: MinimumSnapConfig, waypoints, ns=None, seg_offset=None, v_avg_override=-1.0) -> int:
        wp = _f64(waypoints)
        B, nsu, so, _ = _layout(wp, ns, seg_offset)
        c = cfg.to_c()
        out = C.c_longlong(0)
        self._check(self._L.msnap_sample_bound_host(self._h, C.byref(c), v_avg_override, B, nsu, _ptr(so), _ptr(wp),
                                                    C.byref(out)))
        return int(out.value)

    def generate_batch(self, cfg: MinimumSnapConfig, waypoints, ns=None, seg_offset=None,
                       sample_distance_override=-1.0, v_avg_override=-1.0, capacity: Optional[int] = None,
                       out: Optional[dict] = None, stats: bool = True, outputs: str = "all") -> BatchResult:
        """B independent GenerateTrajectoryMatrix calls in one launch sequence; host arrays in, host arrays out.
        ``capacity`` rows are reserved for the samples (default: the exact-safe upper bound from
        ``msnap_sample_bound``); if it is too small MsnapError(ERR_CAPACITY) is raised with the exact layout
        available in ``.sample_offset`` of the exception's ``partial`` attribute.
        ``outputs``: "all" = every array include/msnap.h offers; "samples" = what the reference's
        GenerateTrajectoryMatrix returns -- the sampled rows (with their CSR offsets and the per-trajectory flags): the
        optional pointers are passed as NULL, nothing else is computed for or copied to the host."""
        if outputs not in ("all", "samples"):
            raise ValueError("outputs must be 'all' or 'samples'")
        wp = _f64(waypoints)
        B, nsu, so, n_seg = _layout(wp, ns, seg_offset)
        c = cfg.to_c()
        if capacity is None:
            capacity = self.sample_bound(cfg, wp, ns=ns, seg_offset=seg_offset, v_avg_override=v_avg_override)
        m = 2 * int(cfg.order)
        o = out or {}
        full = outputs == "all"
        times = o.get("times", np.empty(n_seg)) if full else None
        coeff = o.get("coeff", np.empty((n_seg, 3, m))) if full else None
        max_dev = o.get("max_dev", np.empty(B)) if full else None
        iters = o.get("iters", np.empty(B, dtype=np.int32)) if full else None
        vw_final = o.get("vw_final", np.empty(B)) if full else None
        best_s = o.get("best_s", np.zeros(n_seg, dtype=np.int32)) if full else None
        sample_offset = o.get("sample_offset", np.empty(B + 1, dtype=np.int64))
        samples = o.get("samples", np.empty((max(int(capacity), 1), 3)))
        stats = o.get("stats", np.empty((B, 2))) if (stats and full) else None   # (the reference only prints them, ms.cpp:194)
        flags = o.get("flags", np.zeros(B, dtype=np.uint32))
        rc = self._L.msnap_generate_batch_host(
            self._h, C.byref(c), float(sample_distance_override), float(v_avg_override), B, nsu, _ptr(so), _ptr(wp),
            _ptr(times), _ptr(coeff), _ptr(max_dev), _ptr(iters), _ptr(vw_final), _ptr(best_s), int(capacity),
            _ptr(sample_offset), _ptr(samples), _ptr(stats), _ptr(flags))
        seg_off = so if so is not None else np.arange(B + 1, dtype=np.int64) * nsu
        rows = int(min(sample_offset[B], capacity)) if rc in (_lib.OK, _lib.ERR_CAPACITY) else 0
        res = BatchResult(seg_off, times, coeff, max_dev, iters, vw_final, best_s, sample_offset, samples[:rows], stats,
                          flags)
        if rc != _lib.OK:
            err = MsnapError(rc, self._L.msnap_last_error(self._h).decode())
            err.partial = res
            raise err
        return res

    # ------------------------------------------------------------------ batched, device-resident (torch tensors)
    def generate_batch_dev(self, cfg: MinimumSnapConfig, waypoints, sample_offset, samples, ns=None, seg_offset=None,
                           sample_distance_override=-1.0, v_avg_override=-1.0, times=None, coeff=None, max_dev=None,
                           iters=None, vw_final=None, best_s=None, stats=None, flags=None):
        """Enqueue one batched generate on the handle's stream.  Every array argument is a CUDA ``torch.Tensor`` on
        this handle's device (fp64 / int64 / int32 / uint32-as-int32 as in include/msnap.h); nothing is copied and
        the host is not synchronised."""
        B = int(sample_offset.numel()) - 1
        c = cfg.to_c()

        def dp(t):
            return None if t is None else int(t.data_ptr())

        rc = self._L.msnap_generate_batch_dev(
            self._h, C.byref(c), float(sample_distance_override), float(v_avg_override), B, int(ns or 0),
            dp(seg_offset), dp(waypoints), dp(times), dp(coeff), dp(max_dev), dp(iters), dp(vw_final), dp(best_s),
            int(samples.shape[0]), dp(sample_offset), dp(samples), dp(stats), dp(flags))
        self._check(rc)

    def solve_qp_batch_dev(self, order, waypoints, times, coeff, B, ns=None, seg_offset=None, vel=None, acc=None,
                           path_weight=0.0, vel_zero_weight=0.0, max_dev=None, best_s=None, flags=None):
        def dp(t):
            return None if t is None else int(t.data_ptr())

        rc = self._L.msnap_solve_qp_batch_dev(
            self._h, int(order), float(path_weight), float(vel_zero_weight), int(B), int(ns or 0), dp(seg_offset),
            dp(waypoints), dp(vel), dp(acc), dp(times), dp(coeff), dp(max_dev), dp(best_s), dp(flags))
        self._check(rc)

    def sample_bound_dev(self, cfg: MinimumSnapConfig, waypoints, B, rows_out, ns=None, seg_offset=None,
                         v_avg_override=-1.0):
        c = cfg.to_c()
        rc = self._L.msnap_sample_bound_dev(
            self._h, C.byref(c), float(v_avg_override), int(B), int(ns or 0),
            None if seg_offset is None else int(seg_offset.data_ptr()), int(waypoints.data_ptr()),
            int(rows_out.data_ptr()))
        self._check(rc)


@dataclass
class BezierConfig:
    """struct BezierConfig, bezier.hpp:91-96."""

    min_radius: float = 1.0


class Bezier:
    """GPU-backed stand-in for math_util::Bezier (bezier.hpp:98-120): SetConfig + GenerateTrajectoryMatrix."""

    def __init__(self, tool: Optional[TrajectoryGeneratorTool] = None, device: int = 0):
        self._tool = tool or TrajectoryGeneratorTool(device)
        self._config = BezierConfig()

    def SetConfig(self, config: BezierConfig):
        self._config = config

    def GenerateTrajectoryMatrix(self, Path, yaml_path: str = "", sample_distance_override=-1.0, v_avg_override=-1.0):
        """bezier.cpp:127-189.  Path (n,3) -> sampled points (M,3); fewer than 2 rows -> an empty (0,3) matrix
        (bezier.cpp:129-131).  yaml_path and v_avg_override are unused, as in the reference."""
        Path = np.asarray(Path, dtype=np.float64)
        if Path.ndim != 2 or Path.shape[0] < 2:
            return np.zeros((0, 3))
        wp = _f64(Path[:, :3])
        _, rows, _ = self._tool.bezier_generate_batch(wp, ns=wp.shape[0] - 1, sample_distance_override=sample_distance_override,
                                                      min_radius=self._config.min_radius)
        return rows
