// msnap_geo.cuh -- batched WGS84 <-> ENU maps (sm_100a, fp64): the step every sampled point takes right after the
// minimum-snap sampler in the reference (enuToWGS84_Batch at uavPathPlanning.cpp:3699, 3806, 4909-4913) and every
// waypoint takes right before it (wgs84ToENU_Batch at cpp:2640, 3217).
//
// Mathematical contract, in the reference's operation order (the library is compiled with --fmad=false, so
// a*b + c*d + e*f rounds exactly as the reference's unfused x86-64 code does):
//   calcN, deg2rad, rad2deg, WGS84_A, WGS84_E2                /root/reference/uavPathPlanning.hpp:134-173
//   wgs84ToECEF                                               /root/reference/uavPathPlanning.cpp:894-910
//   ecefToWGS84: start value + <= 10 fixed-point steps,
//                stop when |lat_new - lat| < 1e-12            /root/reference/uavPathPlanning.cpp:926-968
//   ecefToENU / enuToECEF (rotation by the reference point)   /root/reference/uavPathPlanning.cpp:971-1044
//   wgs84ToENU / enuToWGS84                                   /root/reference/uavPathPlanning.cpp:1047-1083
// The only arithmetic that differs from the reference's is inside sin/cos/atan2 (CUDA's fp64 libm, <= 2 ulp, against
// glibc's) -- a few 1e-16 rad, i.e. below 1e-8 m; sqrt and division are IEEE-exact on both sides; pow(x, 3) is
// replaced by an fma-compensated cube that is correctly rounded except in ~1e-16 of the cases, as glibc's pow is.
// That statement-by-statement form is k_enu_to_wgs84<true> (msnap_set_geo_exact_trig); the default,
// k_enu_to_wgs84<false>, runs the same iteration on direction vectors instead of angles (geo_ecef_to_wgs84_fast: same
// start value, same step, same stopping rule, ~4x fewer instructions).
// The quantities that depend on the reference point only (its ECEF position and the two rotation matrices, which
// the reference recomputes for every point) are computed ONCE per call on the host with the host's libm
// (geo_make_frame), i.e. bit-identically to the reference, and passed to the kernels by value.
//
// Work decomposition: one point per lane; a warp moves 32 consecutive rows (768 contiguous bytes) per trip through
// shared memory, so global loads and stores are fully coalesced although a row is 24 bytes; in place is allowed
// (out == in); the next trip's rows are fetched into registers before the current trip's arithmetic.  ENU -> WGS84 is
// bound by the FP64 pipe, not by HBM (measured, B200: statement-by-statement form 1 950 issued instructions per row at
// the usual 4 fixed-point steps -- 7 atan2, 6 sincos, 6 sqrt, ~13 divisions -- 15 G rows/s; direction form ~430 issued,
// ~225 of them FP64, 56 G rows/s = 41 % of the HBM peak); WGS84 -> ENU runs at 77 % of the HBM peak (106 G rows/s) with the
// in-kernel sincos.  48 bytes of traffic per row either way.
#ifndef MSNAP_GEO_CUH
#define MSNAP_GEO_CUH

#include <cmath>

#include "msnap_geo_atan.h"
#include "msnap_geo_sincos.h"

namespace msnap {

constexpr double GEO_A = 6378137.0;           // WGS84_A,  hpp:134
constexpr double GEO_E2 = 0.006694379990141;  // WGS84_E2, hpp:135
constexpr double GEO_PI = 3.14159265358979323846;
constexpr int GEO_BLOCK = 256;
constexpr int GEO_MAX_STEPS = 10;             // cpp:936
constexpr double GEO_TOL = 1e-12;             // cpp:937

// Everything that depends only on the reference point (lon, lat, alt).
struct GeoFrame {
    double ref_ecef[3];
    double R[9];     // ECEF delta -> ENU   (computeENURotationMatrix, cpp:971-994), row-major
    double Rinv[9];  // ENU -> ECEF delta   (computeENURotationMatrixInverse, cpp:998-1020), row-major
};

// Host side, host libm: the same statements the reference executes per point for its reference argument.
inline void geo_make_frame(const double ref_lla[3], GeoFrame &f) {
    const double lat_rad = ref_lla[1] * GEO_PI / 180.0, lon_rad = ref_lla[0] * GEO_PI / 180.0;  // deg2rad, hpp:166-168
    {                                                                                           // wgs84ToECEF, cpp:894-910
        const double sl = std::sin(lat_rad);
        const double N = GEO_A / std::sqrt(1.0 - GEO_E2 * sl * sl);
        const double cos_lat = std::cos(lat_rad), sin_lat = std::sin(lat_rad);
        const double cos_lon = std::cos(lon_rad), sin_lon = std::sin(lon_rad);
        f.ref_ecef[0] = (N + ref_lla[2]) * cos_lat * cos_lon;
        f.ref_ecef[1] = (N + ref_lla[2]) * cos_lat * sin_lon;
        f.ref_ecef[2] = (N * (1 - GEO_E2) + ref_lla[2]) * sin_lat;
    }
    const double cos_lat = std::cos(lat_rad), sin_lat = std::sin(lat_rad);
    const double cos_lon = std::cos(lon_rad), sin_lon = std::sin(lon_rad);
    const double R[9] = {-sin_lon,           cos_lon,            0.0,
                         -sin_lat * cos_lon, -sin_lat * sin_lon, cos_lat,
                         cos_lat * cos_lon,  cos_lat * sin_lon,  sin_lat};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            f.R[3 * i + j] = R[3 * i + j];
            f.Rinv[3 * j + i] = R[3 * i + j];  // the reference writes the transpose out entry by entry: same values
        }
}

// x^3 rounded like a correctly-rounded pow(x, 3): the two products' rounding errors are recovered with fma.
__device__ __forceinline__ double geo_cube(double x) {
    const double s = x * x, es = fma(x, x, -s);
    const double t = s * x, et = fma(s, x, -t);
    return t + (et + es * x);
}

__device__ __forceinline__ double geo_calcN(double sin_lat) {  // hpp:139-142
    return GEO_A / sqrt(1.0 - GEO_E2 * sin_lat * sin_lat);
}

// ecefToWGS84, cpp:926-968.  Returns the number of fixed-point steps taken.
__device__ __forceinline__ int geo_ecef_to_wgs84(double x, double y, double z, double &lon_deg, double &lat_deg,
                                                  double &alt_out) {
    const double p = sqrt(x * x + y * y);
    const double theta = atan2(z * GEO_A, p * GEO_A * (1 - GEO_E2));
    double st, ct;
    sincos(theta, &st, &ct);
    double lat = atan2(z + GEO_E2 * GEO_A * (1 - GEO_E2) * geo_cube(st) / (1 - GEO_E2), p - GEO_E2 * GEO_A * geo_cube(ct));
    int steps = 0;
    double sl, cl;
    sincos(lat, &sl, &cl);
#pragma unroll 1
    for (int i = 0; i < GEO_MAX_STEPS; ++i) {
        const double N = geo_calcN(sl);
        const double alt = p / cl - N;
        const double lat_new = atan2(z, p * (1 - GEO_E2 * N / (N + alt)));
        const bool done = fabs(lat_new - lat) < GEO_TOL;
        lat = lat_new;
        sincos(lat, &sl, &cl);  // needed by the next step, or by the final N / alt below
        ++steps;
        if (done) break;
    }
    const double lon = atan2(y, x);
    const double N = geo_calcN(sl);
    const double alt = p < 1e-12 ? fabs(z) - GEO_A * sqrt(1 - GEO_E2) : p / cl - N;  // cpp:956-960
    lat_deg = lat * 180.0 / GEO_PI;  // rad2deg, hpp:171-173
    lon_deg = lon * 180.0 / GEO_PI;
    alt_out = alt;
    return steps;
}

// The same map without trigonometry inside the loop (default).  A latitude is carried as an unnormalised direction
// d = (S, C), lat = atan2(S, C).  With R = |d| and W = sqrt(C^2 + (1 - e2) S^2) one has cos(lat)/sqrt(1 - e2 sin^2(lat))
// = C / W, and the reference's step  lat_new = atan2(z, p (1 - e2 N / (N + alt)))  with  N + alt = p / cos(lat)
// (cpp:940-944) is  lat_new = atan2(z, p - e2 a C / W) = atan2(z W, p W - e2 a C):  one sqrt and five multiply-adds per
// step, no division, no sin/cos/atan2.  The stopping rule |lat_new - lat| < 1e-12 (cpp:946) is evaluated on the angle
// between the two directions, cross / dot; start value, step limit and final formulas are the reference's.  The angle
// is formed once, at the end.  Directions are rescaled by 2^-23 per step (exact) so that ten steps cannot overflow.
// Differences to geo_ecef_to_wgs84: rounding only (a few 1e-16 rad), plus the stopping decision where |d lat| falls
// within ~1e-4 relative of the threshold (the reference's own subtraction is that noisy there); one step more or less
// moves the result by < 1e-14 rad.
__constant__ double GEO_ATAN_C[GEO_ATAN_N] = GEO_ATAN_COEFFS;

// ---- lean fp64 primitives for the direction form (device only) ------------------------------------------------
// CUDA's sqrt / division / atan2 are IEEE-exact resp. <= 2 ulp but carry slow paths and cost 23 / 20 / 128 issued
// instructions each (ncu, this kernel); these cost 7-11 / 6-8 / ~45 and are accurate to ~1 ulp on the ranges used.
__device__ __forceinline__ double geo_rcp(double x) {  // 1/x, |x| normal: MUFU.RCP64H seed + two Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double geo_div(double n, double d) {  // n/d with one residual correction
    const double r = geo_rcp(d), q = n * r;
    return fma(fma(-q, d, n), r, q);
}
__device__ __forceinline__ double geo_rsqrt(double q) {  // 1/sqrt(q), q > 0 normal: MUFU.RSQ64H seed + one cubic step
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    const double e = fma(-(q * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}
__device__ __forceinline__ double geo_sqrt_pos(double q, double y) {  // sqrt(q) from y = geo_rsqrt(q)
    const double s = q * y;
    return fma(fma(-s, s, q), 0.5 * y, s);
}
// atan2 for finite arguments: t = min/max in [0, 1], atan(t) = t + t^3 P(t^2) (GEO_ATAN_N = 21 coefficients from gen_geo_atan.py:
// Chebyshev interpolation in 60-digit arithmetic, approximation error 3e-17), quadrant fix-up with pi split into two doubles.  atan2(0, 0) = 0.
__device__ __forceinline__ double geo_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const bool sw = ay > ax, xneg = x < 0.0;
    const double mx = sw ? ay : ax, mn = sw ? ax : ay;
    const double t = mn * geo_rcp(mx), u = t * t;  // <= 1.5 ulp: plenty for a result that is rounded to ~1 ulp of pi anyway
    double pl = GEO_ATAN_C[GEO_ATAN_N - 1];
#pragma unroll
    for (int i = GEO_ATAN_N - 2; i >= 0; --i) pl = fma(pl, u, GEO_ATAN_C[i]);  // DFMA with a constant-bank operand
    double a = fma(t * u, pl, t);
    // octant fix-up  a | pi/2 - a | pi/2 + a | pi - a  as  (off_hi + (+-a)) + off_lo
    const double off_hi = sw ? 1.57079632679489655800e+00 : (xneg ? 3.14159265358979311600e+00 : 0.0);
    const double off_lo = sw ? 6.12323399573676603587e-17 : (xneg ? 1.22464679914735320717e-16 : 0.0);
    a = (off_hi + (sw != xneg ? -a : a)) + off_lo;
    if (mx == 0.0) a = 0.0;
    return copysign(a, y);
}
// sin and cos of an angle in radians as it comes out of deg2rad (|x| of a few): Cody-Waite reduction by pi/2 in three
// pieces (k * piece exact for |k| < 2^20), then the two polynomials of msnap_geo_sincos.h on |r| <= pi/4 and the
// quadrant fix-up; ~1 ulp.  Far outside that range (|x| > 1e5, or NaN) libm's sincos with its full reduction takes over.
__constant__ double GEO_SIN_C[GEO_SIN_N] = GEO_SIN_COEFFS;
__constant__ double GEO_COS_C[GEO_COS_N] = GEO_COS_COEFFS;
__device__ __forceinline__ void geo_sincos(double x, double &sn_out, double &cs_out) {
    if (!(fabs(x) <= 1.0e5)) {
        sincos(x, &sn_out, &cs_out);
        return;
    }
    const double kf = rint(x * 0.63661977236758134308);  // 2 / pi
    const int k = (int)kf;
    double r = fma(-kf, GEO_PIO2_1, x);
    r = fma(-kf, GEO_PIO2_2, r);
    r = fma(-kf, GEO_PIO2_3, r);
    const double u = r * r;
    double ps = GEO_SIN_C[GEO_SIN_N - 1], pc = GEO_COS_C[GEO_COS_N - 1];
#pragma unroll
    for (int i = GEO_SIN_N - 2; i >= 0; --i) ps = fma(ps, u, GEO_SIN_C[i]);
#pragma unroll
    for (int i = GEO_COS_N - 2; i >= 0; --i) pc = fma(pc, u, GEO_COS_C[i]);
    const double sn = fma(r * u, ps, r), cs = fma(u * u, pc, fma(-0.5, u, 1.0));
    // x = r + k pi/2:  k mod 4 = 0: (sn, cs)   1: (cs, -sn)   2: (-sn, -cs)   3: (-cs, sn); signs flipped in the sign bit
    const double s = (k & 1) ? cs : sn, c = (k & 1) ? sn : cs;
    sn_out = __hiloint2double(__double2hiint(s) ^ ((k & 2) << 30), __double2loint(s));
    cs_out = __hiloint2double(__double2hiint(c) ^ (((k + 1) & 2) << 30), __double2loint(c));
}

__device__ __forceinline__ double geo_deg2rad(double deg) {  // (deg * pi) / 180, hpp:166-168, division by residual fix-up
    constexpr double INV_180 = 1.0 / 180.0;
    const double n = deg * GEO_PI, q = n * INV_180;
    return fma(fma(-q, 180.0, n), INV_180, q);
}
__device__ __forceinline__ double geo_rad2deg(double rad) {  // (rad * 180) / pi, hpp:171-173, division by residual fix-up
    constexpr double INV_PI = 0.318309886183790671538;
    const double n = rad * 180.0, q = n * INV_PI;
    return fma(fma(-q, GEO_PI, n), INV_PI, q);
}

__device__ __forceinline__ int geo_ecef_to_wgs84_fast(double x, double y, double z, double &lon_deg, double &lat_deg,
                                                       double &alt_out) {
    constexpr double E2A = GEO_E2 * GEO_A, OME2 = 1.0 - GEO_E2, SCALE = 1.1920928955078125e-07;  // 2^-23
    const double p2 = fma(x, x, y * y);
    const double p = p2 > 0.0 ? geo_sqrt_pos(p2, geo_rsqrt(p2)) : 0.0;
    // theta = atan2(z a, p a (1 - e2)): only sin(theta), cos(theta) are used (cpp:929-932)
    const double u = z, v = p * OME2;  // the common factor a of both arguments does not change the direction
    const double ih = geo_rsqrt(fma(u, u, v * v));
    const double st = u * ih, ct = v * ih;
    double S = z + E2A * (st * st * st);            // z + e2 a (1-e2) sin^3 / (1-e2), cpp:931
    double C = fma(-E2A, ct * ct * ct, p);          // p - e2 a cos^3,                  cpp:932
    int steps = 0;
    const double zs = z * SCALE, ps = p * SCALE;  // the per-step rescaling by 2^-23 folded into the constants (exact)
    constexpr double E2AS = E2A * SCALE;
#pragma unroll
    for (int i = 0; i < GEO_MAX_STEPS; ++i) {
        const double q = fma(C, C, OME2 * S * S);
        const double W = q * geo_rsqrt(q);
        const double Sn = zs * W, Cn = fma(ps, W, -E2AS * C);
        const double cross = fma(Sn, C, -(Cn * S)), dot = fma(Sn, S, Cn * C);
        const bool done = fabs(cross) < GEO_TOL * dot;
        S = Sn;
        C = Cn;
        steps = i + 1;
        if (done) break;
    }
    const double ir = geo_rsqrt(fma(S, S, C * C));
    const double sl = S * ir, cl = C * ir;
    const double N = GEO_A * geo_rsqrt(fma(-GEO_E2 * sl, sl, 1.0));
    const double alt = p < 1e-12 ? fabs(z) - GEO_A * sqrt(OME2) : geo_div(p, cl) - N;  // cpp:956-960
    double lat = geo_atan2(S, C);
    if (p == 0.0) lat = NAN;  // the reference's step evaluates 0 * inf there (cpp:944)
    lat_deg = geo_rad2deg(lat);
    lon_deg = geo_rad2deg(geo_atan2(y, x));
    alt_out = alt;
    return steps;
}

// Warp-cooperative move of 32 rows x 3 doubles between global and shared memory (three coalesced 256-byte accesses).
// The loads of the NEXT trip are issued into registers before the current trip's arithmetic (geo_rows_fetch) and
// only parked in shared memory when the trip starts (geo_rows_park), so no warp waits on HBM.
struct GeoRows {
    double v[3];
};
__device__ __forceinline__ GeoRows geo_rows_fetch(const double *g, long long row0, long long n, int lane) {
    GeoRows r;
    const long long base = 3 * row0, end = 3 * n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const long long i = base + 32 * k + lane;
        r.v[k] = i < end ? g[i] : 0.0;
    }
    return r;
}
__device__ __forceinline__ void geo_rows_park(const GeoRows &r, double *sm, int lane) {
#pragma unroll
    for (int k = 0; k < 3; ++k) sm[32 * k + lane] = r.v[k];
    __syncwarp();
}
__device__ __forceinline__ void geo_rows_out(double *__restrict__ g, long long row0, long long n, const double *sm, int lane) {
    __syncwarp();
    const long long base = 3 * row0, end = 3 * n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const long long i = base + 32 * k + lane;
        if (i < end) g[i] = sm[32 * k + lane];
    }
    __syncwarp();
}

// enuToWGS84_Batch (cpp:1098-1108): rows [east, north, up] -> rows [lon_deg, lat_deg, alt_m].
// n_dev != nullptr: the row count is read from device memory (the sampler's sample_offset[B]) and clamped to n_cap.
template <bool TRIG>
__global__ void __launch_bounds__(GEO_BLOCK) k_enu_to_wgs84(GeoFrame f, long long n_cap, const long long *__restrict__ n_dev,
                                                            const double *enu, double *lla, int *__restrict__ steps_out) {
    __shared__ double sm_all[GEO_BLOCK / 32][96];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *sm = sm_all[w];
    long long n = n_cap;
    if (n_dev) {
        const long long m = *n_dev;
        n = m < n_cap ? m : n_cap;
    }
    const long long warps = (long long)gridDim.x * (GEO_BLOCK / 32);
    long long row0 = ((long long)blockIdx.x * (GEO_BLOCK / 32) + w) * 32;
    GeoRows next = geo_rows_fetch(enu, row0, n, lane);
    for (; row0 < n; row0 += warps * 32) {
        geo_rows_park(next, sm, lane);
        // in place: the next trip's rows belong to this warp alone, nobody has written them yet
        next = geo_rows_fetch(enu, row0 + warps * 32, n, lane);
        const double e = sm[3 * lane], no = sm[3 * lane + 1], u = sm[3 * lane + 2];
        // enuToECEF (cpp:1035-1044) and the shift by the reference point (cpp:1076-1079)
        double X, Y, Z;
        if (TRIG) {  // the reference's unfused products and sums
            const double dx = f.Rinv[0] * e + f.Rinv[1] * no + f.Rinv[2] * u;
            const double dy = f.Rinv[3] * e + f.Rinv[4] * no + f.Rinv[5] * u;
            const double dz = f.Rinv[6] * e + f.Rinv[7] * no + f.Rinv[8] * u;
            X = f.ref_ecef[0] + dx, Y = f.ref_ecef[1] + dy, Z = f.ref_ecef[2] + dz;
        } else {
            X = f.ref_ecef[0] + fma(f.Rinv[2], u, fma(f.Rinv[1], no, f.Rinv[0] * e));  // small terms first, then the shift
            Y = f.ref_ecef[1] + fma(f.Rinv[5], u, fma(f.Rinv[4], no, f.Rinv[3] * e));
            Z = f.ref_ecef[2] + fma(f.Rinv[8], u, fma(f.Rinv[7], no, f.Rinv[6] * e));
        }
        double lon, lat, alt;
        const int steps = TRIG ? geo_ecef_to_wgs84(X, Y, Z, lon, lat, alt) : geo_ecef_to_wgs84_fast(X, Y, Z, lon, lat, alt);
        __syncwarp();
        sm[3 * lane] = lon;
        sm[3 * lane + 1] = lat;
        sm[3 * lane + 2] = alt;
        if (steps_out && row0 + lane < n) steps_out[row0 + lane] = steps;
        geo_rows_out(lla, row0, n, sm, lane);
    }
}

// wgs84ToENU_Batch (cpp:1085-1095): rows [lon_deg, lat_deg, alt_m] -> rows [east, north, up].
// TRIG = true: libm sincos and the reference's a / sqrt(..) (msnap_set_geo_exact_trig); false (default): geo_sincos and
// a * rsqrt(..), ~1 ulp each (1e-9 m).
template <bool TRIG>
__global__ void __launch_bounds__(GEO_BLOCK) k_wgs84_to_enu(GeoFrame f, long long n, const double *lla, double *enu) {
    __shared__ double sm_all[GEO_BLOCK / 32][96];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *sm = sm_all[w];
    const long long warps = (long long)gridDim.x * (GEO_BLOCK / 32);
    long long row0 = ((long long)blockIdx.x * (GEO_BLOCK / 32) + w) * 32;
    GeoRows next = geo_rows_fetch(lla, row0, n, lane);
    for (; row0 < n; row0 += warps * 32) {
        geo_rows_park(next, sm, lane);
        next = geo_rows_fetch(lla, row0 + warps * 32, n, lane);
        const double lon_deg = sm[3 * lane], lat_deg = sm[3 * lane + 1], h = sm[3 * lane + 2];
        // wgs84ToECEF (cpp:894-910)
        const double lat_rad = TRIG ? lat_deg * GEO_PI / 180.0 : geo_deg2rad(lat_deg);
        const double lon_rad = TRIG ? lon_deg * GEO_PI / 180.0 : geo_deg2rad(lon_deg);
        double sin_lat, cos_lat, sin_lon, cos_lon;
        double N;
        if (TRIG) {
            sincos(lat_rad, &sin_lat, &cos_lat);
            sincos(lon_rad, &sin_lon, &cos_lon);
            N = geo_calcN(sin_lat);
        } else {
            geo_sincos(lat_rad, sin_lat, cos_lat);
            geo_sincos(lon_rad, sin_lon, cos_lon);
            N = GEO_A * geo_rsqrt(fma(-GEO_E2 * sin_lat, sin_lat, 1.0));
        }
        const double x = (N + h) * cos_lat * cos_lon;
        const double y = (N + h) * cos_lat * sin_lon;
        const double z = (N * (1 - GEO_E2) + h) * sin_lat;
        // delta (cpp:1053-1056) and ecefToENU (cpp:1023-1032)
        const double dx = x - f.ref_ecef[0], dy = y - f.ref_ecef[1], dz = z - f.ref_ecef[2];
        __syncwarp();
        if (TRIG) {  // the reference's unfused products and sums (cpp:1028-1030)
            sm[3 * lane] = f.R[0] * dx + f.R[1] * dy + f.R[2] * dz;
            sm[3 * lane + 1] = f.R[3] * dx + f.R[4] * dy + f.R[5] * dz;
            sm[3 * lane + 2] = f.R[6] * dx + f.R[7] * dy + f.R[8] * dz;
        } else {
            sm[3 * lane] = fma(f.R[2], dz, fma(f.R[1], dy, f.R[0] * dx));
            sm[3 * lane + 1] = fma(f.R[5], dz, fma(f.R[4], dy, f.R[3] * dx));
            sm[3 * lane + 2] = fma(f.R[8], dz, fma(f.R[7], dy, f.R[6] * dx));
        }
        geo_rows_out(enu, row0, n, sm, lane);
    }
}

}  // namespace msnap
#endif  // MSNAP_GEO_CUH
