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
// Work decomposition: k_alt_prep is row-parallel (edge weights of both passes and the follow target: hypot, two divisions
// per row, off the recurrences); k_alt_solve runs one trajectory per lane through pass 1 and the whole active-set loop --
// a dependent chain of n steps per solve (one reciprocal and ~14 multiply-adds per row forward, 2 per row backward), so
// the kernel is latency-bound and is launched with one warp per CTA to spread the chains over all SMs.
#ifndef MSNAP_ALT_CUH
#define MSNAP_ALT_CUH

#include <cmath>

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

// Row-parallel preparation: w1 / w2 = climb weight of the edge (g, g + 1) in pass 1 / pass 2 (0 for the last row of a
// trajectory and for skipped edges), tgt = follow target of pass 1 (NaN where the map has no value).
__global__ void __launch_bounds__(256) k_alt_prep(AltParams p, long long B, const long long *__restrict__ row_offset,
                                                  const double *__restrict__ rows, const double *__restrict__ elev,
                                                  double *__restrict__ w1, double *__restrict__ w2, double *__restrict__ tgt) {
    const long long n = row_offset[B];
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = B;  // trajectory of row g: largest b with row_offset[b] <= g
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (row_offset[mid] <= g) lo = mid;
            else hi = mid;
        }
        const bool last = g + 1 >= row_offset[lo + 1];
        double a1 = 0.0, a2 = 0.0;
        if (!last) {
            const double dist = hypot(rows[3 * (g + 1)] - rows[3 * g], rows[3 * (g + 1) + 1] - rows[3 * g + 1]);
            if (dist > 1e-9) {  // cpp:1655, 1765
                const double d1 = dist * p.max_climb_rate, d2 = dist * (p.max_climb_rate * 0.5);
                if (p.max_climb_rate > 0.0 && d1 > 1e-12) a1 = 1.0 / (d1 * d1);
                if (p.max_climb_rate * 0.5 > 0.0 && d2 > 1e-12) a2 = 1.0 / (d2 * d2);
            }
        }
        w1[g] = a1;
        w2[g] = a2;
        const double el = elev ? elev[g] : NAN;
        tgt[g] = el == el ? fmax(rows[3 * g + 2], el + p.safe_distance) : NAN;  // cpp:1637-1638
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

// One banded LDL' solve of a trajectory (n rows starting at `base`).  diag_extra(k) / rhs(k) supply what the two passes
// add to the shared smoothing + climb + regularisation part.  Leaves z in `out` and returns false on a bad pivot.
template <class Extra, class Rhs, class Sink>
__device__ __forceinline__ bool alt_solve(long long base, long long n, double s, const double *__restrict__ w,
                                          double *__restrict__ l1, double *__restrict__ l2, double *__restrict__ yd,
                                          Extra diag_extra, Rhs rhs, Sink sink) {
    const bool smooth = n >= 3 && s > 0.0;
    double a1 = 0.0, a2 = 0.0, c = 0.0, Dm1 = 0.0, Dm2 = 0.0, ym1 = 0.0, ym2 = 0.0, wm1 = 0.0;
    bool ok = true;
    for (long long k = 0; k < n; ++k) {
        const double wk = w[base + k];
        const int in_m = smooth && k - 1 >= 1 && k - 1 <= n - 2, in_0 = smooth && k >= 1 && k <= n - 2,
                  in_p = smooth && k + 1 >= 1 && k + 1 <= n - 2;
        const double d = s * (double)(in_p + 4 * in_0 + in_m) + (wm1 + wk) + diag_extra(k) + ALT_REG;
        const double e = k + 1 < n ? s * (double)(-2 * (in_0 + in_p)) - wk : 0.0;  // H[k, k+1]
        const double f = k + 2 < n ? s * (double)in_p : 0.0;                       // H[k, k+2]
        const double D = fma(-a2 * a2, Dm2, fma(-a1 * a1, Dm1, d));
        const double y = fma(-a2, ym2, fma(-a1, ym1, rhs(k)));
        ok = ok && D > 0.0 && D < 1e300;
        const double inv = alt_rcp(D);
        const double n1 = fma(-c * Dm1, a1, e) * inv;  // L[k+1, k]
        const double n2 = f * inv;                     // L[k+2, k]
        l1[base + k] = a1;
        l2[base + k] = a2;
        yd[base + k] = y * inv;
        a2 = c;
        a1 = n1;
        c = n2;
        Dm2 = Dm1;
        Dm1 = D;
        ym2 = ym1;
        ym1 = y;
        wm1 = wk;
    }
    double z1 = 0.0, z2 = 0.0, b1 = 0.0, b2 = 0.0, b2n = 0.0;  // z_{k+1}, z_{k+2}; L[k+1,k], L[k+2,k]
    for (long long k = n - 1; k >= 0; --k) {
        const double z = fma(-b2, z2, fma(-b1, z1, yd[base + k]));
        sink(k, z);
        z2 = z1;
        z1 = z;
        b2 = b2n;              // L[(k-1)+2, k-1] = l2[k+1]
        b2n = l2[base + k];    // becomes L[k, k-2], used two rows further down
        b1 = l1[base + k];     // L[k, k-1], used by row k-1
    }
    return ok;
}

// One trajectory per lane: optimizeSegmentAltitudeENU (cpp:1329-1364).  The new heights replace the `up` column of rows.
__global__ void __launch_bounds__(32) k_alt_solve(AltParams p, long long B, const long long *__restrict__ row_offset,
                                                  double *rows, const double *__restrict__ elev,
                                                  const double *__restrict__ w1, const double *__restrict__ w2,
                                                  const double *__restrict__ tgt, double *__restrict__ l1,
                                                  double *__restrict__ l2, double *__restrict__ yd, double *__restrict__ zin,
                                                  double *__restrict__ cur, unsigned char *__restrict__ act,
                                                  double *__restrict__ z_pass1_out, int *__restrict__ solves_out,
                                                  unsigned *__restrict__ flags_out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long base = row_offset[b], n = row_offset[b + 1] - base;
    if (n <= 0) {
        if (solves_out) solves_out[b] = 0;
        if (flags_out) flags_out[b] = 0;
        return;
    }
    // ---- pass 1: optimizeHeights
    bool ok = alt_solve(
        base, n, p.lambda_smooth, w1, l1, l2, yd,
        [&](long long k) { const double t = tgt[base + k]; return t == t ? p.lambda_follow : 0.0; },
        [&](long long k) { const double t = tgt[base + k]; return t == t ? p.lambda_follow * t : 0.0; },
        [&](long long k, double z) {
            const double el = elev ? elev[base + k] : NAN;
            if (el == el && z < el + p.safe_distance) z = el + p.safe_distance;  // cpp:1705-1707
            zin[base + k] = z;
            act[base + k] = 0;
            if (z_pass1_out) z_pass1_out[base + k] = z;
        });
    // ---- pass 2: optimizeHeightsGlobalSmooth with lambda_smooth * 10, max_climb_rate * 0.5 (cpp:1352-1355)
    const double s2 = p.lambda_smooth * 10.0;
    int solves = 0;
    for (int iter = 0; iter < ALT_MAX_ITER; ++iter) {
        bool violation = false;
        ok = alt_solve(
                 base, n, s2, w2, l1, l2, yd,
                 [&](long long k) {
                     double x = 0.0;
                     if (k == 0) x += ALT_FIX_WEIGHT;
                     if (k == n - 1) x += ALT_FIX_WEIGHT;
                     if (k >= 1 && k < n - 1 && act[base + k]) x += ALT_CON_WEIGHT;
                     return x;
                 },
                 [&](long long k) {
                     const double zi = zin[base + k];
                     double x = 0.0;
                     if (k == 0) x += ALT_FIX_WEIGHT * zi;
                     if (k == n - 1) x += ALT_FIX_WEIGHT * zi;
                     if (k >= 1 && k < n - 1 && act[base + k]) x += ALT_CON_WEIGHT * zi;
                     return x;
                 },
                 [&](long long k, double z) {
                     cur[base + k] = z;
                     if (z < zin[base + k] - ALT_VIOLATION && !act[base + k]) {  // cpp:1805-1810
                         act[base + k] = 1;
                         violation = true;
                     }
                 }) &&
             ok;
        ++solves;
        if (!violation) break;
    }
    for (long long k = 0; k < n; ++k) {  // cpp:1817-1821, written back as segment_enu[i].up (cpp:1357-1359)
        const double z = cur[base + k], zi = zin[base + k];
        rows[3 * (base + k) + 2] = z < zi ? zi : z;
    }
    if (solves_out) solves_out[b] = solves;
    if (flags_out) flags_out[b] = ok ? 0u : ALT_FLAG_PIVOT;
}

}  // namespace msnap
#endif  // MSNAP_ALT_CUH
