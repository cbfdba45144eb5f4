"""oracle/msnap_oracle.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement ("port") of the reference's minimum-snap hot path, written to follow
``/root/reference/math_util/minimum_snap.cpp`` statement by statement: the same dense matrices M, C_T, Q, A, V,
the same products in the same left-to-right order, dense inverses by partial-pivot LU (``numpy.linalg.inv`` =
LAPACK getrf/getri; Eigen's ``MatrixXd::inverse()`` is PartialPivLU), the same two-pass path penalty, the same
reweighting loop and the same distance-thresholded sampler.  Each function cites the reference lines it follows.

PINNING.  The reference ships no test, golden file or recorded output for this path, so
this restatement is pinned against outputs of the reference's own unmodified source executed in the build
container (``oracle/_ref`` = minimum_snap.cpp compiled against ``oracle/shim/Eigen/Dense``); those outputs are
committed as ``tests/golden/*.npz`` together with ``tests/golden/make_golden.py``, and
``tests/test_oracle.py`` checks this file against every one of them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class MinimumSnapConfig:
    """minimum_snap.hpp:9-33 (defaults are the struct's; the shipped YAML overrides them)."""

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


def factorial(x: int) -> int:
    """ms.cpp:15-20."""
    fac = 1
    i = x
    while i > 0:
        fac *= i
        i -= 1
    return fac


def _cdiv(a: int, b: int) -> int:
    """C++ int division (truncation toward zero); operands here are non-negative."""
    return int(a / b) if (a < 0) != (b < 0) else a // b


@dataclass
class SolveInfo:
    max_dev: float = 0.0
    best_t: np.ndarray = None          # seg_best_t           (ms.cpp:342, 463)
    best_s: np.ndarray = None          # index s in 0..16 of the arg-max sample (ms.cpp:415-439)
    best_dist2: np.ndarray = None      # seg_best_dist2_before (ms.cpp:343, 464)
    dist2_table: np.ndarray = None     # all 17 squared deviations per segment (for tie classification)
    ratio: np.ndarray = None           # per-segment deviation ratio after the solve (ms.cpp:613-617)


def solve_qp_closed_form(order, Path, Vel, Acc, Time, path_weight=0.0, vel_zero_weight=0.0):
    """TrajectoryGeneratorTool::SolveQPClosedForm, ms.cpp:227-649.

    Path (n,3), Vel/Acc (2,3), Time (ns,).  Returns (PolyCoeff (ns, 3*2*order), SolveInfo)."""
    Path = np.asarray(Path, dtype=np.float64)
    Vel = np.asarray(Vel, dtype=np.float64)
    Acc = np.asarray(Acc, dtype=np.float64)
    Time = np.asarray(Time, dtype=np.float64)
    p_order = 2 * order - 1                       # ms.cpp:237
    p_num1d = p_order + 1                         # ms.cpp:238
    number_segments = Time.shape[0]               # ms.cpp:240
    PolyCoeff = np.zeros((number_segments, 3 * p_num1d))
    number_coefficients = p_num1d * number_segments

    # ---- M, ms.cpp:247-266
    Mb_rows, Mb_cols = order * 2, p_num1d
    M = np.zeros((number_segments * Mb_rows, number_segments * Mb_cols))
    for i in range(number_segments):
        row, col = i * Mb_rows, i * Mb_cols
        sub_M = np.zeros((Mb_rows, Mb_cols))
        for j in range(order):
            for k in range(p_num1d):
                if k < j:
                    continue
                ratio = _cdiv(factorial(k), factorial(k - j))
                sub_M[j, p_num1d - 1 - k] = ratio * math.pow(0.0, k - j)
                sub_M[j + order, p_num1d - 1 - k] = ratio * math.pow(Time[i], k - j)
        M[row:row + Mb_rows, col:col + Mb_cols] = sub_M

    # ---- C_T, ms.cpp:269-310
    number_valid_variables = (number_segments + 1) * order
    number_fixed_variables = 2 * order + (number_segments - 1)
    n_free = number_valid_variables - number_fixed_variables
    C_T = np.zeros((number_coefficients, number_valid_variables))
    for i in range(number_coefficients):
        if i < order:
            C_T[i, i] = 1
            continue
        if i >= number_coefficients - order:
            delta_index = i - (number_coefficients - order)
            C_T[i, number_fixed_variables - order + delta_index] = 1
            continue
        if (i % order == 0) and (i // order % 2 == 1):
            C_T[i, i // (2 * order) + order] = 1
            continue
        if (i % order == 0) and (i // order % 2 == 0):
            C_T[i, i // (2 * order) + order - 1] = 1
            continue
        if (i % order != 0) and (i // order % 2 == 1):
            t0 = i // (2 * order) * (2 * order) + order
            t1 = i // (2 * order) * (order - 1) + i - t0 - 1
            C_T[i, number_fixed_variables + t1] = 1
            continue
        if (i % order != 0) and (i // order % 2 == 0):
            t0 = (i - order) // (2 * order) * (2 * order) + order
            t1 = (i - order) // (2 * order) * (order - 1) + (i - order) - t0 - 1
            C_T[i, number_fixed_variables + t1] = 1
            continue

    # ---- Q, ms.cpp:313-330 (integer arithmetic before the pow, as written)
    Q = np.zeros((number_coefficients, number_coefficients))
    for k in range(number_segments):
        sub_Q = np.zeros((p_num1d, p_num1d))
        for i in range(p_order + 1):
            for l in range(p_order + 1):
                if p_num1d - i <= order or p_num1d - l <= order:
                    continue
                ai = _cdiv(factorial(p_order - i), factorial(p_order - order - i))
                al = _cdiv(factorial(p_order - l), factorial(p_order - order - l))
                e = p_order - i + p_order - l - (2 * order - 1)
                sub_Q[i, l] = _cdiv(ai * al, e) * math.pow(Time[k], e)
        r = k * p_num1d
        Q[r:r + p_num1d, r:r + p_num1d] = sub_Q

    f_coeff = [np.zeros(number_coefficients) for _ in range(3)]   # ms.cpp:332-334
    Q_original = Q.copy()                                          # ms.cpp:337
    A = np.zeros((number_coefficients, number_coefficients))      # ms.cpp:340
    seg_best_t = np.zeros(number_segments)                        # ms.cpp:342
    seg_best_s = np.zeros(number_segments, dtype=np.int64)
    seg_best_dist2_before = np.zeros(number_segments)             # ms.cpp:343
    dist2_table = np.zeros((number_segments, 17))

    # every textual M.inverse() / M.transpose().inverse() in the reference re-factorises the same matrix;
    # the result is identical each time, so it is computed once here.
    M_inv = np.linalg.inv(M)
    Mt_inv = np.linalg.inv(M.T.copy())

    def d_selected_for(axis):
        """ms.cpp:357-389 and 524-562 (identical loops)."""
        d_selected = np.zeros(number_valid_variables)
        for i in range(number_coefficients):
            if i == 0:
                d_selected[i] = Path[0, axis]
                continue
            if i == 1 and order >= 2:
                d_selected[i] = Vel[0, axis]
                continue
            if i == 2 and order >= 3:
                d_selected[i] = Acc[0, axis]
                continue
            if i == number_coefficients - order + 2 and order >= 3:
                d_selected[number_fixed_variables - order + 2] = Acc[1, axis]
                continue
            if i == number_coefficients - order + 1 and order >= 2:
                d_selected[number_fixed_variables - order + 1] = Vel[1, axis]
                continue
            if i == number_coefficients - order:
                d_selected[number_fixed_variables - order] = Path[number_segments, axis]
                continue
            if (i % order == 0) and (i // order % 2 == 0):
                d_selected[i // (2 * order) + order - 1] = Path[i // (2 * order), axis]
                continue
        return d_selected

    nf = number_fixed_variables
    if path_weight > 0.0:                                          # ms.cpp:347
        # 1) initial solve with Q_original only, ms.cpp:349-405
        R_tmp = C_T.T @ Mt_inv @ Q_original @ M_inv @ C_T          # ms.cpp:350
        P0 = []
        for axis in range(3):
            d_selected = d_selected_for(axis)
            R_PP_tmp = R_tmp[nf:, nf:]
            d_F_tmp = d_selected[:nf]
            R_FP_tmp = R_tmp[:nf, nf:]
            if n_free > 0:
                d_opt_tmp = ((-np.linalg.inv(R_PP_tmp)) @ R_FP_tmp.T) @ d_F_tmp   # ms.cpp:398
                d_selected[nf:] = d_opt_tmp
            d_tmp = C_T @ d_selected
            P0.append(M_inv @ d_tmp)                               # ms.cpp:402-404
        # 2) per-segment worst-deviation search, ms.cpp:408-465
        nsamples = 16
        for k in range(number_segments):
            T = Time[k]
            best_t, best_dist2, best_s = 0.0, -1.0, 0
            row = k * p_num1d
            for s in range(nsamples + 1):
                tt = T * float(s) / float(nsamples)
                phi = np.array([math.pow(tt, p_order - i) for i in range(p_num1d)])
                x = phi.dot(P0[0][row:row + p_num1d])
                y = phi.dot(P0[1][row:row + p_num1d])
                z = phi.dot(P0[2][row:row + p_num1d])
                L = Path[k] + (tt / T) * (Path[k + 1] - Path[k])
                d = np.array([x, y, z]) - L
                dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
                dist2_table[k, s] = dist2
                if dist2 > best_dist2:
                    best_dist2, best_t, best_s = dist2, tt, s
            phi_best = np.array([math.pow(best_t, p_order - i) for i in range(p_num1d)])
            A[row:row + p_num1d, row:row + p_num1d] = np.outer(phi_best, phi_best)   # ms.cpp:445-446
            Lbest = Path[k] + (best_t / T) * (Path[k + 1] - Path[k])                # ms.cpp:451
            for i in range(p_num1d):
                for axis in range(3):
                    f_coeff[axis][row + i] = -2.0 * (phi_best[i] * Lbest[axis]) * path_weight   # ms.cpp:452-460
            seg_best_t[k] = best_t
            seg_best_s[k] = best_s
            seg_best_dist2_before[k] = best_dist2
        Q = Q + path_weight * A                                    # ms.cpp:468

    # velocity soft penalty, ms.cpp:474-509
    if vel_zero_weight > 0.0:
        V = np.zeros((number_coefficients, number_coefficients))

        def eval_phi_dot(t):
            phi_d = np.zeros(p_num1d)
            for i in range(p_num1d):
                power = p_order - i - 1
                if power < 0:
                    phi_d[i] = 0.0
                elif power == 0:
                    phi_d[i] = float(p_order - i)
                else:
                    phi_d[i] = float(p_order - i) * math.pow(t, power)
            return phi_d

        for k in range(number_segments):
            T = Time[k]
            row = k * p_num1d
            s0 = eval_phi_dot(0.0)
            V[row:row + p_num1d, row:row + p_num1d] += np.outer(s0, s0)
            s1 = eval_phi_dot(T)
            V[row:row + p_num1d, row:row + p_num1d] += np.outer(s1, s1)
        Q = Q + vel_zero_weight * V                                # ms.cpp:506

    R = C_T.T @ Mt_inv @ Q @ M_inv @ C_T                            # ms.cpp:511
    f_valid = [np.zeros(number_valid_variables) for _ in range(3)]
    if path_weight > 0.0:                                           # ms.cpp:517-522
        for axis in range(3):
            f_valid[axis] = (C_T.T @ Mt_inv) @ f_coeff[axis]

    P = []
    for axis in range(3):                                           # ms.cpp:524-592
        d_selected = d_selected_for(axis)
        R_PP = R[nf:, nf:]
        d_F = d_selected[:nf]
        R_FP = R[:nf, nf:]
        f_P = f_valid[axis][nf:]
        if n_free > 0:
            d_optimal = (-np.linalg.inv(R_PP)) @ (R_FP.T @ d_F + f_P)   # ms.cpp:579 (f not halved, as written)
            d_selected[nf:] = d_optimal
        d = C_T @ d_selected
        P.append(M_inv @ d)                                         # ms.cpp:584-591

    # post-solve deviation at the recorded t*, ms.cpp:594-624
    current_max_dev = 0.0
    ratios = np.zeros(number_segments)
    for k in range(number_segments):
        best_t = seg_best_t[k]
        phi_best = np.array([math.pow(best_t, p_order - i) for i in range(p_num1d)])
        row = k * p_num1d
        fin = np.array([phi_best.dot(P[a][row:row + p_num1d]) for a in range(3)])
        T = Time[k]
        Lbest = Path[k] + (best_t / T) * (Path[k + 1] - Path[k])
        dd = fin - Lbest
        dist_after = math.sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2])
        sv = Path[k + 1] - Path[k]
        seg_len = math.sqrt(sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2])
        ratio = dist_after / seg_len if seg_len > 1e-6 else 0.0
        ratios[k] = ratio
        if ratio > current_max_dev:
            current_max_dev = ratio

    for i in range(number_segments):                                # ms.cpp:626-646
        for j in range(3):
            PolyCoeff[i, j * p_num1d:(j + 1) * p_num1d] = P[j][p_num1d * i:p_num1d * (i + 1)]

    info = SolveInfo(current_max_dev, seg_best_t, seg_best_s, seg_best_dist2_before, dist2_table, ratios)
    return PolyCoeff, info


def allocate_time(Path, V_avg, min_time_s):
    """ms.cpp:63-72.  Plain IEEE mul/add (the reference as shipped is built without FMA contraction)."""
    Path = np.asarray(Path, dtype=np.float64)
    ns = Path.shape[0] - 1
    Time = np.zeros(ns)
    for i in range(ns):
        dx = Path[i + 1, 0] - Path[i, 0]
        dy = Path[i + 1, 1] - Path[i, 1]
        dz = Path[i + 1, 2] - Path[i, 2]
        ln = math.sqrt(dx * dx + dy * dy + dz * dz)
        t = (ln / V_avg) if V_avg > 1e-6 else min_time_s
        if t < min_time_s:
            t = min_time_s
        Time[i] = t
    return Time


@dataclass
class GenerateInfo:
    Time: np.ndarray = None
    PolyCoeff: np.ndarray = None
    max_dev: float = 0.0
    iters: int = 0
    vw_final: float = 0.0
    max_climb_rate: float = 0.0
    min_turn_radius: float = 1.0e12
    seg_counts: np.ndarray = None      # samples recorded inside each segment's loop (ms.cpp:140-152)
    n_candidates: np.ndarray = None
    last_solve: SolveInfo = None
    min_accept_margin: float = math.inf   # min |seg_len - sample_distance| over all candidates (tie classification)


def eval_poly_at(PolyCoeff, p_num1d, seg, t):
    """ms.cpp:104-117: sum of c * pow(t, exp), highest power first, accumulated left to right."""
    pt = np.zeros(3)
    for dim in range(3):
        val = 0.0
        for k in range(p_num1d):
            c = PolyCoeff[seg, dim * p_num1d + k]
            val += c * math.pow(t, p_num1d - 1 - k)
        pt[dim] = val
    return pt


def _norm3(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def sample_polynomials(polyCoeff, Time, order, sample_distance):
    """The sampler of GenerateTrajectoryMatrix, ms.cpp:97-161, on given coefficients (PolyCoeff matrix, ns x 3*2*order)
    and segment times.  Returns (list of points, per-segment accepted counts, per-segment candidate counts, min
    |seg_len - sample_distance| over all candidates)."""
    num_segments = len(Time)
    p_num1d = 2 * order
    dt_default = 0.1                                  # ms.cpp:100
    samples = []
    has_last = False
    seg_counts = np.zeros(num_segments, dtype=np.int64)
    n_cand = np.zeros(num_segments, dtype=np.int64)
    margin = math.inf
    for seg in range(num_segments):                   # ms.cpp:123-161
        T = Time[seg]
        dt = dt_default
        if dt > T / 10.0:
            dt = T / 10.0
        t0_pt = eval_poly_at(polyCoeff, p_num1d, seg, 0.0)
        if not has_last:
            samples.append(t0_pt)
            has_last = True
        prev_pt = t0_pt
        t = dt
        while t <= T + 1e-12:
            tt = min(t, T)
            cur_pt = eval_poly_at(polyCoeff, p_num1d, seg, tt)
            seg_len = _norm3(cur_pt - prev_pt)
            margin = min(margin, abs(seg_len - sample_distance))
            n_cand[seg] += 1
            if seg_len >= sample_distance:
                prev_pt = cur_pt
                samples.append(cur_pt)
                seg_counts[seg] += 1
            t += dt
        if seg == num_segments - 1:
            endpt = eval_poly_at(polyCoeff, p_num1d, seg, T)
            if len(samples) == 0 or _norm3(samples[-1] - endpt) > 1e-6:
                samples.append(endpt)
    return samples, seg_counts, n_cand, margin


def generate_trajectory_matrix(Path, cfg: MinimumSnapConfig, sample_distance_override=-1.0, v_avg_override=-1.0):
    """TrajectoryGeneratorTool::GenerateTrajectoryMatrix, ms.cpp:22-206.  Returns (samples (S,3), GenerateInfo);
    an input with fewer than 2 rows or 3 columns returns an empty (0,0) matrix like ms.cpp:54-57."""
    Path = np.asarray(Path, dtype=np.float64)
    order = cfg.order
    V_avg = cfg.V_avg
    min_time_s = cfg.min_time_s
    sample_distance = cfg.sample_distance
    Vel = np.zeros((2, 3))
    Acc = np.zeros((2, 3))
    Vel[0], Vel[1] = cfg.start_vel, cfg.end_vel
    Acc[0], Acc[1] = cfg.start_acc, cfg.end_acc
    path_weight = cfg.path_weight
    vel_zero_weight = cfg.vel_zero_weight
    if sample_distance_override > 0.0:
        sample_distance = sample_distance_override
    if v_avg_override > 0.0:
        V_avg = v_avg_override
    if Path.ndim != 2 or Path.shape[0] < 2 or Path.shape[1] < 3:
        return np.zeros((0, 0)), GenerateInfo()

    num_segments = Path.shape[0] - 1
    Time = allocate_time(Path, V_avg, min_time_s)

    # reweighting loop, ms.cpp:76-90
    max_iter, it = 10, 0
    while True:
        polyCoeff, sinfo = solve_qp_closed_form(order, Path, Vel, Acc, Time, path_weight, vel_zero_weight)
        if sinfo.max_dev > 0.2 and it < max_iter:
            vel_zero_weight = 0.01 if vel_zero_weight < 1e-6 else vel_zero_weight * 2.0
            it += 1
        else:
            break

    samples, seg_counts, n_cand, margin = sample_polynomials(polyCoeff, Time, order, sample_distance)

    # climb / turn statistics, ms.cpp:163-195 (printed by the reference, returned here)
    max_climb_rate, min_turn_radius = 0.0, 1.0e12
    for i in range(len(samples) - 1):
        dx = samples[i + 1][0] - samples[i][0]
        dy = samples[i + 1][1] - samples[i][1]
        dz = abs(samples[i + 1][2] - samples[i][2])
        h = math.sqrt(dx * dx + dy * dy)
        if h > 1e-6:
            rate = dz / h
            if rate > max_climb_rate:
                max_climb_rate = rate
        if i > 0:
            p0, p1, p2 = samples[i - 1], samples[i], samples[i + 1]
            a, b, c = _norm3(p1 - p0), _norm3(p2 - p1), _norm3(p2 - p0)
            area = 0.5 * _norm3(np.cross(p1 - p0, p2 - p0))
            if area > 1e-8:
                Rr = (a * b * c) / (4.0 * area)
                if Rr < min_turn_radius:
                    min_turn_radius = Rr

    out = np.array(samples).reshape(-1, 3)
    info = GenerateInfo(Time, polyCoeff, sinfo.max_dev, it, vel_zero_weight, max_climb_rate, min_turn_radius,
                        seg_counts, n_cand, sinfo, margin)
    return out, info
