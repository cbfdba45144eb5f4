// msnap_alt.cuh -- batched altitude optimisation of sampled trajectories (sm_100a, fp64): the step the reference runs on
// the minimum-snap sampler's output before converting it to WGS84 (getPlan, uavPathPlanning.cpp:3712-3729 ->
// runAltitudeOptimization cpp:1535-1573 -> optimizeSegmentAltitudeENU cpp:1329-1364).  SURVEY.md section 8f rank 2.
//
// Mathematical contract (one independent problem per trajectory of n rows [east, north, up]):
//   pass 1  optimizeHeights (cpp:1575-1712):  H z = b with
//             H = lambda_smooth * L'L  (L = second difference on interior rows, cpp:1588-1604)
//               + lambda_follow * diag(has_elev)                                   (cpp:1607-1645)
//               + sum over edges w_i (e_i - e_{i+1})(e_i - e_{i+1})',  w_i = 1 / (dist_i * max_climb_rate)^2, edges with
//                 dist_i <= 1e-9 or dist_i * max_climb_rate <= 1e-12 skipped      (cpp:1649-1665)
//               + 1e-8 I                                                           (cpp:1668-1670)
//             b_i = lambda_follow * max(up_i, elev_i + safe_distance) where the map has a value, else 0
//           then z_i = max(z_i, elev_i + safe_distance) where the map has a value   (cpp:1684-1709)
//   pass 2  optimizeHeightsGlobalSmooth (cpp:1714-1827) with lambda_smooth * 10 and max_climb_rate * 0.5 (cpp:1352-1355):
//           at most 10 solves of  (lambda L'L + climb + 1e10 (e_0 e_0' + e_{n-1} e_{n-1}') + 1e8 diag(active) + 1e-8 I) z = b,
//           b = 1e10 z1 at both ends + 1e8 z1 on active rows; a row becomes active when z_i < z1_i - 1e-3; stop when no row
//           was added; finally z = max(z, z1).
// The reference assembles H as an Eigen sparse matrix and factors it with SimplicialLDLT (fill-reducing ordering); H is
// SPD pentadiagonal, so here each trajectory runs a banded LDL' recurrence (half-bandwidth 2) in natural order: same
// system, same solution, different rounding (compare DESIGN.md section 10 for the bar).
//
// Work decomposition: k_alt_prep + k_alt_ends are row- / trajectory-parallel (edge weights of both passes and the follow
// target: hypot, two divisions per row, off the recurrences); k_alt_solve_pair (default: a lane pair per trajectory,
// two-sided elimination) or k_alt_solve (one lane per trajectory) runs pass 1 and the whole active-set loop -- dependent
// chains of n (or n/2) steps per solve, one reciprocal and ~14 multiply-adds per row forward, 2 per row backward -- so the
// kernels are latency-bound and are launched with one warp per CTA to spread the chains over all SMs; k_alt_finish
// (row-parallel) writes the heights back into the rows.
#ifndef MSNAP_ALT_CUH
#define MSNAP_ALT_CUH

#include <cuda_pipeline.h>

#include <cmath>
#include <type_traits>

namespace msnap {

struct AltParams {  // struct AltitudeParams, uavPathPlanning.hpp:415-421
    double lambda_smooth, lambda_follow, max_climb_rate, uav_R, safe_distance;
};

constexpr double ALT_REG = 1e-8;          // cpp:1669, 1791
constexpr double ALT_FIX_WEIGHT = 1e10;   // cpp:1779
constexpr double ALT_CON_WEIGHT = 1e8;    // cpp:1787
constexpr double ALT_VIOLATION = 1e-3;    // cpp:1805
constexpr int ALT_MAX_ITER = 10;          // cpp:1733
constexpr unsigned ALT_FLAG_PIVOT = 1u;   // a non-positive or non-finite pivot appeared (the reference's "decomposition failed")
constexpr unsigned ALT_FLAG_TRUNCATED = 2u;  // the trajectory's rows do not fit n_rows_cap: skipped, rows untouched
constexpr unsigned ALT_FLAG_PASS2 = 4u;   // the failure happened in pass 2: the pass-1 heights were kept (cpp:1356)
// per-trajectory outcome handed from the solve kernel to k_alt_mark / k_alt_finish
constexpr int ALT_ST_OK = 0, ALT_ST_KEEP_INPUT = 1, ALT_ST_KEEP_PASS1 = 2;

// ElevationCostMap::getCostAt (elevation_cost_map.cpp:373-380): nearest cell of a row-major float grid with a top-left
// origin; NaN where the reference returns false.
__global__ void __launch_bounds__(256) k_cost_lookup(const float *__restrict__ grid, int width, int height, double resolution,
                                                     double origin_x, double origin_y, long long n_cap,
                                                     const long long *__restrict__ n_dev, const double *__restrict__ rows,
                                                     double *__restrict__ elev) {
    long long n = n_cap;
    if (n_dev) n = *n_dev < n_cap ? *n_dev : n_cap;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const double x = rows[3 * g], y = rows[3 * g + 1];
        const double fc = floor((x - origin_x) / resolution), fr = floor((origin_y - y) / resolution);
        double v = NAN;
        if (fc >= 0.0 && fc < (double)width && fr >= 0.0 && fr < (double)height)
            v = (double)grid[(long long)fr * width + (long long)fc];
        elev[g] = v;
    }
}

// Row-parallel preparation: w1 / w2 = climb weight of the edge (g, g + 1) in pass 1 / pass 2 (0 for skipped edges), tgt =
// follow target of pass 1 (NaN where the map has no value), act = 0.  Every row is treated as if its successor belonged to
// the same trajectory; k_alt_ends (one thread per trajectory, launched right after) zeroes the weights of each
// trajectory's last row, so no row has to search for its trajectory.
__global__ void __launch_bounds__(256) k_alt_prep(AltParams p, long long B, const long long *__restrict__ row_offset,
                                                  const double *__restrict__ rows, const double *__restrict__ elev,
                                                  double *__restrict__ w1, double *__restrict__ w2, double *__restrict__ tgt,
                                                  double *__restrict__ act, long long n_cap) {
    const long long n = row_offset[B] < n_cap ? row_offset[B] : n_cap;  // rows beyond the caller's buffers do not exist
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const double x0 = rows[3 * g], y0 = rows[3 * g + 1], up = rows[3 * g + 2];
        double a1 = 0.0, a2 = 0.0;
        if (g + 1 < n) {
            const double dist = hypot(rows[3 * (g + 1)] - x0, rows[3 * (g + 1) + 1] - y0);
            if (dist > 1e-9) {  // cpp:1655, 1765
                const double d1 = dist * p.max_climb_rate, d2 = dist * (p.max_climb_rate * 0.5);
                if (p.max_climb_rate > 0.0 && d1 > 1e-12) a1 = 1.0 / (d1 * d1);
                if (p.max_climb_rate * 0.5 > 0.0 && d2 > 1e-12) a2 = 1.0 / (d2 * d2);
            }
        }
        w1[g] = a1;
        w2[g] = a2;
        act[g] = 0.0;
        const double el = elev ? elev[g] : NAN;
        tgt[g] = el == el ? fmax(up, el + p.safe_distance) : NAN;  // cpp:1637-1638
    }
}
__global__ void __launch_bounds__(256) k_alt_ends(long long B, const long long *__restrict__ row_offset, double *__restrict__ w1,
                                                  double *__restrict__ w2, long long n_cap) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long first = row_offset[b], end = row_offset[b + 1];
    if (end > first && end <= n_cap) {  // the last row of a trajectory has no edge to a successor
        w1[end - 1] = 0.0;
        w2[end - 1] = 0.0;
    }
}

__device__ __forceinline__ double alt_rcp(double x) {  // 1/x for a positive normal pivot: MUFU seed + two Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// ---- k_alt_solve: one trajectory per lane, 32 trajectories per warp, one warp per CTA -------------------------------
// The kernel is a set of dependent chains, so what counts is the latency of one row step and nothing may wait on memory
// inside it.  Each lane therefore streams the rows of ITS trajectory through shared memory in chunks of 32 rows with
// cp.async (8 bytes per copy, consecutive addresses: every 32-byte sector is fetched from L2 once and served from L1 three
// more times), double buffered: the copies of the next chunk are issued before the current chunk's 32 recurrence steps and
// waited for after them.  A lane's tile row is private to it ([field][lane][row], padded pitch => no bank conflicts), so no
// cross-lane synchronisation or index exchange is needed; results leave by plain stores, which nothing waits for.
constexpr int ALT_CHUNK = 32;
constexpr int ALT_PITCH = ALT_CHUNK + 1;  // lane t reads word t * 33 + j: conflict-free
constexpr int ALT_TILES = 5;              // fields a sweep stages at most (backward sweep of pass 2)
constexpr int ALT_TILE_WORDS = 32 * ALT_PITCH;
constexpr size_t ALT_SMEM_BYTES = 2 * ALT_TILES * ALT_TILE_WORDS * sizeof(double);

// One sweep over this lane's rows, ascending (forward elimination) or descending (back substitution).  in[f] is staged
// into tile f; `row(k, j, T)` performs row k with T(f, j) = staged value of field f.  Lanes with active == false idle.
template <int NIN, bool DESC, class Row>
__device__ __forceinline__ void alt_sweep(double *sm, const double *const (&in)[NIN], long long base, int n, int nmax, bool active,
                                          int lane, Row row) {
    if (nmax <= 0) return;
    const int n_chunks = (nmax + ALT_CHUNK - 1) / ALT_CHUNK;
    const int n_eff = active ? n : 0;
    auto chunk_k0 = [&](int c) { return (DESC ? n_chunks - 1 - c : c) * ALT_CHUNK; };
    auto issue = [&](int c, int buf) {
        const int k0 = chunk_k0(c);
        double *mine = sm + (size_t)buf * ALT_TILES * ALT_TILE_WORDS + lane * ALT_PITCH;
        const int cnt = n_eff - k0 < ALT_CHUNK ? n_eff - k0 : ALT_CHUNK;  // rows of this chunk this lane owns (may be <= 0)
#pragma unroll
        for (int f = 0; f < NIN; ++f) {
            const double *src = in[f] + base + k0;
            double *dst = mine + f * ALT_TILE_WORDS;
#pragma unroll 8
            for (int j = 0; j < ALT_CHUNK; ++j)
                if (j < cnt) __pipeline_memcpy_async(dst + j, src + j, sizeof(double));
        }
        __pipeline_commit();
    };
    issue(0, 0);
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        const bool more = c + 1 < n_chunks;
        if (more) issue(c + 1, buf ^ 1);
        if (more) __pipeline_wait_prior(1);
        else __pipeline_wait_prior(0);
        const double *mine = sm + (size_t)buf * ALT_TILES * ALT_TILE_WORDS + lane * ALT_PITCH;
        auto T = [&](int f, int j) { return mine[f * ALT_TILE_WORDS + j]; };
        const int k0 = chunk_k0(c);
        const int cnt = n_eff - k0 < ALT_CHUNK ? n_eff - k0 : ALT_CHUNK;
        if (DESC) {
            for (int j = cnt - 1; j >= 0; --j) row(k0 + j, j, T);
        } else {
            for (int j = 0; j < cnt; ++j) row(k0 + j, j, T);
        }
        __syncwarp();
    }
}

template <int NN>
struct AltPtrs {
    static constexpr int N = NN;
    const double *p[NN];
};
template <int NN>
struct AltInts {
    static constexpr int N = NN;
    int p[NN];
};

struct AltFwd {  // state a lane carries from row to row of the forward sweep
    double a1 = 0.0, a2 = 0.0, c = 0.0, Dm1 = 0.0, Dm2 = 0.0, ym1 = 0.0, ym2 = 0.0, wm1 = 0.0;
    double sm = 0.0, s0 = 0.0;  // lambda * [row k-1 / row k is an interior row of the smoothing stencil] (row k+1 is tested anew)
    bool ok = true;
};
struct AltBwd {  // ... and of the backward sweep: z_{k+1}, z_{k+2}, L[k+1,k], L[k+2,k], L[k+1,k-1]
    double z1 = 0.0, z2 = 0.0, b1 = 0.0, b2 = 0.0, b2n = 0.0;
};

// Row k of the banded LDL' factorisation fused with the forward substitution.  wk = climb weight of edge (k, k+1),
// extra / rhs = what the pass adds to the diagonal / right-hand side.  Returns (L[k,k-1], L[k,k-2], y_k / D_k).
__device__ __forceinline__ void alt_fwd_row(AltFwd &f, int k, int n, double s, bool smooth, double wk, double extra,
                                            double rhs, double &o_l1, double &o_l2, double &o_yd) {
    // smoothing stencil (cpp:1588-1604): H[k,k] = s (in_p + 4 in_0 + in_m), H[k,k+1] = -2 s (in_0 + in_p), H[k,k+2] = s in_p,
    // in_x = 1 iff row k-1 / k / k+1 is interior (1 <= row <= n-2)
    const double sm = f.sm, s0 = f.s0, sp = (smooth && k + 1 <= n - 2) ? s : 0.0;
    f.sm = s0;
    f.s0 = sp;
    const double d = fma(4.0, s0, sp + sm) + (f.wm1 + wk) + extra + ALT_REG;
    const double e = -2.0 * (s0 + sp) - wk;  // H[k, k+1]; the last row has s0 = sp = 0 and wk = 0 (k_alt_prep)
    const double h = sp;                     // H[k, k+2]; sp != 0 implies k + 2 < n
    const double D = fma(-f.a2 * f.a2, f.Dm2, fma(-f.a1 * f.a1, f.Dm1, d));
    const double y = fma(-f.a2, f.ym2, fma(-f.a1, f.ym1, rhs));
    f.ok = f.ok && D > 0.0 && D < 1e300;
    const double inv = alt_rcp(D);
    const double n1 = fma(-f.c * f.Dm1, f.a1, e) * inv;  // L[k+1, k]
    const double n2 = h * inv;                           // L[k+2, k]
    o_l1 = f.a1;
    o_l2 = f.a2;
    o_yd = y * inv;
    f.a2 = f.c;
    f.a1 = n1;
    f.c = n2;
    f.Dm2 = f.Dm1;
    f.Dm1 = D;
    f.ym2 = f.ym1;
    f.ym1 = y;
    f.wm1 = wk;
}
// Row k of the backward substitution: z_k = y_k/D_k - L[k+1,k] z_{k+1} - L[k+2,k] z_{k+2}.
__device__ __forceinline__ double alt_bwd_row(AltBwd &r, double l1k, double l2k, double ydk) {
    const double z = fma(-r.b2, r.z2, fma(-r.b1, r.z1, ydk));
    r.z2 = r.z1;
    r.z1 = z;
    r.b2 = r.b2n;  // L[k+1, k-1] = l2[k+1]
    r.b2n = l2k;   // L[k, k-2], needed two rows further down
    r.b1 = l1k;    // L[k, k-1], needed by row k-1
    return z;
}

// optimizeSegmentAltitudeENU (cpp:1329-1364) for 32 trajectories per warp.  The heights of both passes stay in the
// scratch arrays zin (pass 1, clamped) and cur (last solve of pass 2); k_alt_finish writes max(cur, zin) back to the rows.
__global__ void __launch_bounds__(32) k_alt_solve(AltParams p, long long B, const long long *__restrict__ row_offset,
                                                  const double *elev, const double *w1, const double *w2, const double *tgt,
                                                  double *l1, double *l2, double *yd, double *zin, double *cur, double *act,
                                                  double *z_pass1_out, int *__restrict__ solves_out,
                                                  unsigned *__restrict__ flags_out, long long n_cap, int *__restrict__ st_out) {
    extern __shared__ double alt_sm[];
    const int lane = threadIdx.x;
    const long long b = (long long)blockIdx.x * 32 + lane;
    const long long base = b < B ? row_offset[b] : 0;
    const bool truncated = b < B && row_offset[b + 1] > n_cap;  // rows missing from the caller's buffers: skip the trajectory
    const int n = (b < B && !truncated) ? (int)(row_offset[b + 1] - base) : 0;
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    bool ok = true, ok2 = true;

    // ---- pass 1: optimizeHeights (cpp:1575-1712)
    {
        const double s = p.lambda_smooth, lf = p.lambda_follow, safe = p.safe_distance;
        const bool smooth = n >= 3 && s > 0.0;
        AltFwd f;
        const double *const in_f[2] = {w1, tgt};
        alt_sweep<2, false>(alt_sm, in_f, base, n, nmax, n > 0, lane, [&](int k, int j, auto &T) {
            const double t = T(1, j);
            const bool has = t == t;
            double o1, o2, o3;
            alt_fwd_row(f, k, n, s, smooth, T(0, j), has ? lf : 0.0, has ? lf * t : 0.0, o1, o2, o3);
            l1[base + k] = o1;
            l2[base + k] = o2;
            yd[base + k] = o3;
        });
        ok = f.ok;
        AltBwd r;
        const bool have_elev = elev != nullptr;
        const double *const in_b[4] = {l1, l2, yd, have_elev ? elev : yd};
        alt_sweep<4, true>(alt_sm, in_b, base, n, nmax, n > 0, lane, [&](int k, int j, auto &T) {
            double z = alt_bwd_row(r, T(0, j), T(1, j), T(2, j));
            const double el = have_elev ? T(3, j) : NAN;
            if (el == el && z < el + safe) z = el + safe;  // cpp:1705-1707
            zin[base + k] = z;
            if (z_pass1_out) z_pass1_out[base + k] = z;
        });
    }
    // ---- pass 2: optimizeHeightsGlobalSmooth with lambda_smooth * 10, max_climb_rate * 0.5 (cpp:1352-1355, 1714-1827)
    const double s2 = p.lambda_smooth * 10.0;
    const bool smooth2 = n >= 3 && s2 > 0.0;
    int solves = 0;
    bool running = n > 0;  // this lane's trajectory still iterates
    for (int iter = 0; iter < ALT_MAX_ITER; ++iter) {
        if (!__any_sync(0xffffffffu, running)) break;
        AltFwd f;
        const double *const in_f[3] = {w2, zin, act};
        alt_sweep<3, false>(alt_sm, in_f, base, n, nmax, running, lane, [&](int k, int j, auto &T) {
            const double zi = T(1, j);
            double x = 0.0;
            if (k == 0) x += ALT_FIX_WEIGHT;                                // cpp:1779-1784
            if (k == n - 1) x += ALT_FIX_WEIGHT;
            if (k >= 1 && k < n - 1 && T(2, j) != 0.0) x += ALT_CON_WEIGHT;  // cpp:1787-1793
            double o1, o2, o3;
            alt_fwd_row(f, k, n, s2, smooth2, T(0, j), x, x * zi, o1, o2, o3);
            l1[base + k] = o1;
            l2[base + k] = o2;
            yd[base + k] = o3;
        });
        if (running) ok2 = ok2 && f.ok;
        AltBwd r;
        bool violation = false;
        const double *const in_b[5] = {l1, l2, yd, zin, act};
        alt_sweep<5, true>(alt_sm, in_b, base, n, nmax, running, lane, [&](int k, int j, auto &T) {
            const double z = alt_bwd_row(r, T(0, j), T(1, j), T(2, j));
            cur[base + k] = z;
            if (z < T(3, j) - ALT_VIOLATION && T(4, j) == 0.0) {  // cpp:1805-1810
                act[base + k] = 1.0;
                violation = true;
            }
        });
        if (running) {
            ++solves;
            if (!violation) running = false;  // converged (cpp:1814)
        }
    }
    if (b < B) {
        if (solves_out) solves_out[b] = solves;
        if (flags_out)
            flags_out[b] = truncated ? ALT_FLAG_TRUNCATED : (!ok ? ALT_FLAG_PIVOT : (!ok2 ? (ALT_FLAG_PIVOT | ALT_FLAG_PASS2) : 0u));
        st_out[b] = (truncated || !ok) ? ALT_ST_KEEP_INPUT : (!ok2 ? ALT_ST_KEEP_PASS1 : ALT_ST_OK);
    }
}

// ---- k_alt_solve_pair: two lanes per trajectory (two-sided / twisted elimination) ----------------------------------
// The chain of a solve is cut in two: lane 0 of a pair eliminates rows 0 .. m-1 downwards, lane 1 rows n-1 .. m+2 upwards
// (the same recurrence on the mirrored system: the smoothing stencil and the penalties are symmetric, the climb weight of
// the edge towards the NEXT row is w[k] going down and w[k-1] going up).  Both stop in front of the two middle rows m, m+1,
// whose 2 x 2 Schur complement collects the contributions of both sides; the pair exchanges six numbers by shuffle, both
// lanes solve the 2 x 2 system (symmetric expressions: bit-identical on both lanes) and substitute back outwards.  Half
// the dependent steps per solve, twice the warps.  Trajectories shorter than ALT_TWIST_MIN rows run on lane 0 alone.
constexpr int ALT_TWIST_MIN = 8;

struct AltSpan {
    long long base;
    int n, side, cnt;  // cnt = local rows this lane walks (its side's rows + its middle row)
    bool twisted;
    __device__ __forceinline__ int row(int i) const { return side ? n - 1 - i : i; }
};

// alt_sweep for a lane of a pair: local indices 0 .. S.cnt-1 (ascending, or descending for the back substitution) map to
// rows S.row(i); field f is read at row + shift[f] on the upward side (the climb weights).
template <int NIN, bool DESC, class Row>
__device__ __forceinline__ void alt_sweep_pair(double *sm, const double *const (&in)[NIN], const int (&shift)[NIN], const AltSpan &S,
                                               int cmax, bool active, int lane, Row row) {
    if (cmax <= 0) return;
    const int n_chunks = (cmax + ALT_CHUNK - 1) / ALT_CHUNK;
    const int len = active ? S.cnt : 0;
    auto chunk_i0 = [&](int c) { return (DESC ? n_chunks - 1 - c : c) * ALT_CHUNK; };
    // Both sides copy their chunk in ASCENDING address order (constant offsets in the unrolled copy loop); the upward side
    // then reads its tile back to front: local row i0 + j sits in slot cnt - 1 - j.
    auto issue = [&](int c, int buf) {
        const int i0 = chunk_i0(c);
        double *mine = sm + (size_t)buf * ALT_TILES * ALT_TILE_WORDS + lane * ALT_PITCH;
        const int cnt = len - i0 < ALT_CHUNK ? len - i0 : ALT_CHUNK;
        const int lo = S.side ? S.row(i0 + (cnt > 0 ? cnt - 1 : 0)) : S.row(i0);  // lowest row of the chunk
#pragma unroll
        for (int f = 0; f < NIN; ++f) {
            const double *src = in[f] + S.base + lo + (S.side ? shift[f] : 0);
            double *dst = mine + f * ALT_TILE_WORDS;
#pragma unroll 8
            for (int j = 0; j < ALT_CHUNK; ++j)
                if (j < cnt) __pipeline_memcpy_async(dst + j, src + j, sizeof(double));
        }
        __pipeline_commit();
    };
    issue(0, 0);
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        const bool more = c + 1 < n_chunks;
        if (more) issue(c + 1, buf ^ 1);
        if (more) __pipeline_wait_prior(1);
        else __pipeline_wait_prior(0);
        const int i0 = chunk_i0(c);
        const int cnt = len - i0 < ALT_CHUNK ? len - i0 : ALT_CHUNK;
        const double *mine = sm + (size_t)buf * ALT_TILES * ALT_TILE_WORDS + lane * ALT_PITCH + (S.side ? cnt - 1 : 0);
        const int dir = S.side ? -1 : 1;  // slot and row step per local row
        const double *q = mine;           // slot of the current local row; T(f, .) = field f of that row (constant offsets)
        auto T = [&](int f, int) { return q[f * ALT_TILE_WORDS]; };
        if (DESC) {
            q = mine + dir * (cnt - 1);
            int k = S.row(i0 + cnt - 1);
#pragma unroll 4
            for (int j = cnt - 1; j >= 0; --j, q -= dir, k -= dir) row(i0 + j, k, j, T);
        } else {
            int k = S.row(i0);
#pragma unroll 4
            for (int j = 0; j < cnt; ++j, q += dir, k += dir) row(i0 + j, k, j, T);
        }
        __syncwarp();
    }
}

// The middle row of a side: its diagonal and right-hand side with this side's eliminations applied, the side's coupling
// term K = L[m+1,m-1] D_{m-1} L[m,m-1] of the off-diagonal entry, and the entry H[m,m+1] itself.  The state is not advanced.
__device__ __forceinline__ void alt_twist_row(const AltFwd &f, int i, int n, double s, bool smooth, double wk, double extra,
                                              double rhs, double &tD, double &tY, double &tK, double &tE) {
    const double sp = (smooth && i + 1 <= n - 2) ? s : 0.0;
    const double d = fma(4.0, f.s0, sp + f.sm) + (f.wm1 + wk) + extra + ALT_REG;
    tE = -2.0 * (f.s0 + sp) - wk;
    tD = fma(-f.a2 * f.a2, f.Dm2, fma(-f.a1 * f.a1, f.Dm1, d));
    tY = fma(-f.a2, f.ym2, fma(-f.a1, f.ym1, rhs));
    tK = f.c * f.Dm1 * f.a1;
}

__global__ void __launch_bounds__(32) k_alt_solve_pair(AltParams p, long long B, const long long *__restrict__ row_offset,
                                                       const double *elev, const double *w1, const double *w2,
                                                       const double *tgt, double *l1, double *l2, double *yd, double *zin,
                                                       double *cur, double *act, double *z_pass1_out,
                                                       int *__restrict__ solves_out, unsigned *__restrict__ flags_out,
                                                       long long n_cap, int *__restrict__ st_out) {
    extern __shared__ double alt_sm[];
    const int lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const long long b = (long long)blockIdx.x * 16 + (lane >> 1);
    AltSpan S;
    S.side = lane & 1;
    S.base = b < B ? row_offset[b] : 0;
    const bool truncated = b < B && row_offset[b + 1] > n_cap;  // rows missing from the caller's buffers: skip the trajectory
    S.n = (b < B && !truncated) ? (int)(row_offset[b + 1] - S.base) : 0;
    S.twisted = S.n >= ALT_TWIST_MIN;
    const int n = S.n, m = (n - 2) / 2;  // middle rows m, m + 1
    S.cnt = S.twisted ? (S.side ? n - m - 1 : m + 1) : (S.side ? 0 : n);
    const int cmax = __reduce_max_sync(FULL, S.cnt);
    const long long base = S.base;
    const int cnt = S.cnt;
    const bool twist = S.twisted;
    // per-trajectory views of the scratch arrays (row k of this trajectory = index k)
    double *const l1_b = l1 + base, *const l2_b = l2 + base, *const yd_b = yd + base, *const zin_b = zin + base,
                  *const cur_b = cur + base, *const act_b = act + base,
                  *const z_pass1_out_b = z_pass1_out ? z_pass1_out + base : nullptr;
    bool ok = true, ok2 = true;
    auto X = [&](double v) { return __shfl_xor_sync(FULL, v, 1); };

    // one solve: forward sweeps of both sides, the middle 2 x 2 system, back substitution outwards.
    //   coef(k, j, T, extra, rhs): what the pass adds to row k's diagonal and right-hand side (T(f, j) = staged field f)
    //   sink(k, j, T, z): what the pass does with z_k (T = the fields staged for the backward sweep)
    auto solve = [&](double s, auto &in_f, auto &sh_f, auto &in_b, auto &sh_b, bool active, auto coef, auto sink) {
        const bool smooth = n >= 3 && s > 0.0;
        AltFwd f;
        double tD = 1.0, tY = 0.0, tK = 0.0, tE = 0.0;
        alt_sweep_pair<std::remove_reference_t<decltype(in_f)>::N, false>(
            alt_sm, in_f.p, sh_f.p, S, cmax, active, lane, [&](int i, int k, int j, auto &T) {
                double extra, rhs;
                coef(k, j, T, extra, rhs);
                if (twist && i == cnt - 1) {
                    alt_twist_row(f, i, n, s, smooth, T(0, j), extra, rhs, tD, tY, tK, tE);
                } else {
                    double o1, o2, o3;
                    alt_fwd_row(f, i, n, s, smooth, T(0, j), extra, rhs, o1, o2, o3);
                    l1_b[k] = o1;
                    l2_b[k] = o2;
                    yd_b[k] = o3;
                }
            });
        AltBwd r;
        double z_mid = 0.0;
        // the partner's middle row and last eliminated row (all lanes take part in the shuffles)
        const double pD = X(tD), pY = X(tY), pK = X(tK), pc = X(f.c), pDm1 = X(f.Dm1), pym1 = X(f.ym1);
        bool good = f.ok;
        if (twist) {
            const double A_self = fma(-pc * pc, pDm1, tD), A_oth = fma(-f.c * f.c, f.Dm1, pD);
            const double A12 = tE - (tK + pK);
            const double g_self = fma(-pc, pym1, tY), g_oth = fma(-f.c, f.ym1, pY);
            const double det = fma(A_self, A_oth, -(A12 * A12));
            good = good && A_self > 0.0 && det > 0.0 && det < 1e300;
            const double idet = alt_rcp(det);
            z_mid = fma(g_self, A_oth, -(g_oth * A12)) * idet;
            r.z1 = z_mid;
            r.z2 = fma(g_oth, A_self, -(g_self * A12)) * idet;
            r.b1 = f.a1;   // L[m, m-1]   (mirrored on the upward side)
            r.b2 = f.c;    // L[m+1, m-1]
            r.b2n = f.a2;  // L[m, m-2]
        }
        alt_sweep_pair<std::remove_reference_t<decltype(in_b)>::N, true>(
            alt_sm, in_b.p, sh_b.p, S, cmax, active, lane, [&](int i, int k, int j, auto &T) {
                const double z = (twist && i == cnt - 1) ? z_mid : alt_bwd_row(r, T(0, j), T(1, j), T(2, j));
                sink(k, j, T, z);
            });
        return good;
    };

    // ---- pass 1: optimizeHeights (cpp:1575-1712)
    {
        const double lf = p.lambda_follow, safe = p.safe_distance;
        const bool have_elev = elev != nullptr;
        const AltPtrs<2> in_f{{w1, tgt}};
        const AltInts<2> sh_f{{-1, 0}};
        const AltPtrs<4> in_b{{l1, l2, yd, have_elev ? elev : yd}};
        const AltInts<4> sh_b{{0, 0, 0, 0}};
        ok = solve(p.lambda_smooth, in_f, sh_f, in_b, sh_b, cnt > 0,
                   [&](int, int j, auto &T, double &extra, double &rhs) {
                       const double t = T(1, j);
                       const bool has = t == t;
                       extra = has ? lf : 0.0;
                       rhs = has ? lf * t : 0.0;
                   },
                   [&](int k, int j, auto &T, double z) {
                       const double el = have_elev ? T(3, j) : NAN;
                       if (el == el && z < el + safe) z = el + safe;  // cpp:1705-1707
                       zin_b[k] = z;
                       if (z_pass1_out_b) z_pass1_out_b[k] = z;
                   });
    }
    // ---- pass 2: optimizeHeightsGlobalSmooth with lambda_smooth * 10, max_climb_rate * 0.5 (cpp:1352-1355, 1714-1827)
    int solves = 0;
    bool running = n > 0;  // the pair's trajectory still iterates (identical on both lanes)
    for (int iter = 0; iter < ALT_MAX_ITER; ++iter) {
        if (!__any_sync(FULL, running)) break;
        bool violation = false;
        const AltPtrs<3> in_f{{w2, zin, act}};
        const AltInts<3> sh_f{{-1, 0, 0}};
        const AltPtrs<5> in_b{{l1, l2, yd, zin, act}};
        const AltInts<5> sh_b{{0, 0, 0, 0, 0}};
        const bool good = solve(p.lambda_smooth * 10.0, in_f, sh_f, in_b, sh_b, running && cnt > 0,
                                [&](int k, int j, auto &T, double &extra, double &rhs) {
                                    double x = 0.0;
                                    if (k == 0) x += ALT_FIX_WEIGHT;                                // cpp:1779-1784
                                    if (k == n - 1) x += ALT_FIX_WEIGHT;
                                    if (k >= 1 && k < n - 1 && T(2, j) != 0.0) x += ALT_CON_WEIGHT;  // cpp:1787-1793
                                    extra = x;
                                    rhs = x * T(1, j);
                                },
                                [&](int k, int j, auto &T, double z) {
                                    cur_b[k] = z;
                                    if (z < T(3, j) - ALT_VIOLATION && T(4, j) == 0.0) {  // cpp:1805-1810
                                        act_b[k] = 1.0;
                                        violation = true;
                                    }
                                });
        const bool partner_violation = __shfl_xor_sync(FULL, violation, 1);  // (no short circuit around a warp collective)
        violation = violation || partner_violation;
        if (running) {
            ok2 = ok2 && good;
            ++solves;
            if (!violation) running = false;  // converged (cpp:1814)
        }
    }
    const bool partner_ok = __shfl_xor_sync(FULL, ok, 1), partner_ok2 = __shfl_xor_sync(FULL, ok2, 1);
    ok = ok && partner_ok;
    ok2 = ok2 && partner_ok2;
    if (b < B && S.side == 0) {
        if (solves_out) solves_out[b] = solves;
        if (flags_out)
            flags_out[b] = truncated ? ALT_FLAG_TRUNCATED : (!ok ? ALT_FLAG_PIVOT : (!ok2 ? (ALT_FLAG_PIVOT | ALT_FLAG_PASS2) : 0u));
        st_out[b] = (truncated || !ok) ? ALT_ST_KEEP_INPUT : (!ok2 ? ALT_ST_KEEP_PASS1 : ALT_ST_OK);
    }
}

// Failure semantics of optimizeSegmentAltitudeENU (cpp:1342-1344, 1356): if pass 1 fails the rows stay untouched, if pass 2
// fails the pass-1 heights are kept.  Thread per trajectory; a trajectory that did not finish both passes poisons its scratch
// heights with NaN (zin and cur, or cur alone), which k_alt_finish reads as "keep".  Failures are rare, so this kernel
// normally only reads B status words.
__global__ void __launch_bounds__(256) k_alt_mark(long long B, const long long *__restrict__ row_offset, long long n_cap,
                                                  const int *__restrict__ st, double *__restrict__ zin, double *__restrict__ cur) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int s = st[b];
    if (s == ALT_ST_OK) return;
    const long long first = row_offset[b];
    long long end = row_offset[b + 1];
    if (end > n_cap) end = n_cap;
    for (long long g = first; g < end; ++g) {
        cur[g] = NAN;
        if (s == ALT_ST_KEEP_INPUT) zin[g] = NAN;
    }
}

// cpp:1817-1821 and the write-back of cpp:1357-1359: up_i = max(z_i, z1_i), row-parallel.  NaN marks from k_alt_mark: no
// pass-1 height => the row keeps its input `up`; no pass-2 height => the pass-1 height.
__global__ void __launch_bounds__(256) k_alt_finish(long long B, const long long *__restrict__ row_offset,
                                                    const double *__restrict__ cur, const double *__restrict__ zin,
                                                    double *__restrict__ rows, long long n_cap) {
    const long long n = row_offset[B] < n_cap ? row_offset[B] : n_cap;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const double z = cur[g], zi = zin[g];
        if (zi == zi) rows[3 * g + 2] = z >= zi ? z : zi;  // (z NaN => zi)
    }
}

}  // namespace msnap
#endif  // MSNAP_ALT_CUH
