// msnap_bezier.cuh -- batched math_util::Bezier::GenerateTrajectoryMatrix (sm_100a, fp64): the reference's alternative
// trajectory generator behind the same Matrix-in / Matrix-out shape as the minimum-snap one (SURVEY.md section 8f rank 4;
// /root/reference/math_util/bezier.cpp:127-189 around Bezier::GeneratePath, bezier.cpp:29-118; selected by
// getPlan(algorithm == "bezier") through UavPathPlanner::Bezier_3D, uavPathPlanning.cpp:3691-3692, 4477-4505).
//
// Contract per trajectory of n waypoints [x, y, z]:
//   heading_i = atan2 of the central difference of the xy waypoints (one-sided at both ends)      bezier.cpp:146-160
//   segment i = cubic Bezier p0 = P_i, p3 = P_{i+1}, p1 = p0 + k d (cos h_i, sin h_i), p2 = p3 - k d (cos h_{i+1}, sin h_{i+1}),
//               d = |P_{i+1} - P_i|_xy, z control points at thirds of the height difference        bezier.cpp:35-101
//   k = 1/3, raised in steps of 0.02 (at most to 0.45) until the curvature at t = 0, 0.5, 1 is <= 1/min_radius -- only when
//               min_radius > 1                                                                     bezier.cpp:46-94
//   samples at t = 0, r, 2r, ... <= 1 with r = resolution / (|p2 - p1|_xy + 2 d / 3), t ACCUMULATED by t += r; every segment
//               after the first drops its t = 0 point                                             bezier.cpp:103-116, 170-174
//   a segment with d < 0.1 m contributes its end waypoint only                                     bezier.cpp:41, 175-179
//
// Two passes around an exclusive scan of the exact row counts (the count of a segment is the number of accumulated t <= 1,
// a rounding-sensitive discrete quantity: both passes run the same `t += r` loop, so layout and rows always agree):
//   k_bezier_prep   thread per segment: headings, control points, k search, r, row count            -> workspace
//   k_bezier_rows   thread per trajectory: start row of every segment, rows of the trajectory       -> scan (msnap_generic.cuh)
//   k_bezier_write  thread per segment: the samples, same arithmetic in the same order as bezier.cpp:110-112
// The library is compiled with --fmad=false and the sample expressions are written with explicit __dmul_rn / __dadd_rn in
// the reference's left-to-right order, so given the same control points the rows are the reference's bits; the control
// points go through atan2 / cos / sin / hypot, where CUDA's and the host's libm may differ by an ulp (1e-12 m).
#ifndef MSNAP_BEZIER_CUH
#define MSNAP_BEZIER_CUH

#include "msnap_generic.cuh"

namespace msnap {

constexpr int BEZ_WS = 8;  // doubles per segment in the workspace: p1.xyz, p2.xyz, r, (count as a double bit pattern)
constexpr unsigned BEZ_FLAG_NONFINITE = 1u;  // == MSNAP_FLAG_NONFINITE: a non-finite waypoint / step made a segment unsampleable
constexpr double BEZ_MIN_STEP = 1.0e-7;      // smallest parameter step the sampler walks (1e7 samples per segment)

struct BezPoint {
    double x, y, z, heading;
};

// heading of waypoint i of a trajectory with n points (bezier.cpp:146-160); p = the trajectory's first waypoint
__device__ __forceinline__ double bez_heading(const double *p, int i, int n) {
    double dx, dy;
    if (i == 0) {
        dx = p[3] - p[0];
        dy = p[4] - p[1];
    } else if (i == n - 1) {
        dx = p[3 * i] - p[3 * (i - 1)];
        dy = p[3 * i + 1] - p[3 * (i - 1) + 1];
    } else {
        dx = p[3 * (i + 1)] - p[3 * (i - 1)];
        dy = p[3 * (i + 1) + 1] - p[3 * (i - 1) + 1];
    }
    return atan2(dy, dx);
}

// Control points of one segment (Bezier::GeneratePath, bezier.cpp:29-101).  Returns false for the d < 0.1 fallback.
__device__ __forceinline__ bool bez_control(const BezPoint &p0, const BezPoint &p3, double min_r, double resolution,
                                            BezPoint &p1, BezPoint &p2, double &step) {
    const double d = hypot(p0.x - p3.x, p0.y - p3.y);
    if (!(d >= 1e-1)) return false;  // (also a NaN distance: the reference's `d < 1e-1` is false there and it would loop on)
    double k = 1.0 / 3.0;
    double s0, c0, s3, c3;
    sincos(p0.heading, &s0, &c0);
    sincos(p3.heading, &s3, &c3);
    auto place = [&](double kk) {
        p1.x = p0.x + c0 * d * kk;
        p1.y = p0.y + s0 * d * kk;
        p1.z = p0.z + (p3.z - p0.z) * 1.0 / 3.0;
        p2.x = p3.x - c3 * d * kk;
        p2.y = p3.y - s3 * d * kk;
        p2.z = p0.z + (p3.z - p0.z) * 2.0 / 3.0;
    };
    for (int iter = 0; iter < 10; ++iter) {
        place(k);
        if (min_r <= 1.0) break;
        bool satisfied = true;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {  // curvature at t = 0, 0.5, 1 (bezier.cpp:60-85)
            const double t = 0.5 * q, it = 1.0 - t;
            const double dx = 3 * it * it * (p1.x - p0.x) + 6 * it * t * (p2.x - p1.x) + 3 * t * t * (p3.x - p2.x);
            const double dy = 3 * it * it * (p1.y - p0.y) + 6 * it * t * (p2.y - p1.y) + 3 * t * t * (p3.y - p2.y);
            const double dz = 3 * it * it * (p1.z - p0.z) + 6 * it * t * (p2.z - p1.z) + 3 * t * t * (p3.z - p2.z);
            const double ddx = 6 * it * (p2.x - 2 * p1.x + p0.x) + 6 * t * (p3.x - 2 * p2.x + p1.x);
            const double ddy = 6 * it * (p2.y - 2 * p1.y + p0.y) + 6 * t * (p3.y - 2 * p2.y + p1.y);
            const double ddz = 6 * it * (p2.z - 2 * p1.z + p0.z) + 6 * t * (p3.z - 2 * p2.z + p1.z);
            const double cx = dy * ddz - dz * ddy, cy = dz * ddx - dx * ddz, cz = dx * ddy - dy * ddx;
            const double cross_norm = sqrt(cx * cx + cy * cy + cz * cz);
            const double vel_norm = sqrt(dx * dx + dy * dy + dz * dz);
            const double vel_norm3 = vel_norm * vel_norm * vel_norm;
            if (vel_norm3 > 1e-6) {
                const double curvature = cross_norm / vel_norm3;
                if (curvature > 1.0 / min_r) {
                    satisfied = false;
                    break;
                }
            }
        }
        if (satisfied) break;
        k += 0.02;
        if (k > 0.45) {
            k = 0.45;
            break;
        }
    }
    place(k);  // bezier.cpp:96-101
    const double dis = hypot(p2.x - p1.x, p2.y - p1.y) + d * 2.0 / 3.0;
    step = resolution / dis;
    return true;
}

// one sample (bezier.cpp:110-112), in the reference's operation order
__device__ __forceinline__ double bez_eval(double t, double it, double a0, double a1, double a2, double a3) {
    const double w0 = __dmul_rn(__dmul_rn(__dmul_rn(it, it), it), a0);
    const double w1 = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(3.0, it), it), t), a1);
    const double w2 = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(3.0, it), t), t), a2);
    const double w3 = __dmul_rn(__dmul_rn(__dmul_rn(t, t), t), a3);
    return __dadd_rn(__dadd_rn(__dadd_rn(w0, w1), w2), w3);
}

template <bool WRITE>
__device__ __forceinline__ int bez_walk(double step, const BezPoint &p0, const BezPoint &p1, const BezPoint &p2,
                                        const BezPoint &p3, bool skip_first, long long row, long long capacity,
                                        double *__restrict__ samples, bool &dropped) {
    int n = 0;
    for (double t = 0.0; t <= 1.0; t = __dadd_rn(t, step)) {
        if (!(skip_first && n == 0)) {
            if (WRITE) {
                if (row < capacity) {
                    const double it = __dsub_rn(1.0, t);
                    samples[3 * row] = bez_eval(t, it, p0.x, p1.x, p2.x, p3.x);
                    samples[3 * row + 1] = bez_eval(t, it, p0.y, p1.y, p2.y, p3.y);
                    samples[3 * row + 2] = bez_eval(t, it, p0.z, p1.z, p2.z, p3.z);
                } else {
                    dropped = true;
                }
                ++row;
            }
        }
        ++n;
    }
    return n - (skip_first ? 1 : 0);
}

// Thread per segment.  ws[g] = {p1.xyz, p2.xyz, step, rows of the segment}; step < 0 marks the end-point-only fallback.
__global__ void __launch_bounds__(128) k_bezier_prep(BatchIdx bi, const double *__restrict__ wp, double resolution, double min_r,
                                                     double *__restrict__ ws, unsigned *__restrict__ flags) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    const double *p = wp + 3 * (g + b);        // waypoint k of trajectory b
    const double *first = p - 3 * k;
    BezPoint p0{p[0], p[1], p[2], bez_heading(first, k, ns + 1)};
    BezPoint p3{p[3], p[4], p[5], bez_heading(first, k + 1, ns + 1)};
    BezPoint p1{}, p2{};
    double step = -1.0;
    int rows = 1;  // fallback: the end waypoint (bezier.cpp:175-179)
    bool bad = false;
    if (bez_control(p0, p3, min_r, resolution, p1, p2, step)) {
        if (step >= BEZ_MIN_STEP) {  // (false for NaN too)
            bool dropped = false;
            rows = bez_walk<false>(step, p0, p1, p2, p3, k > 0, 0, 0, nullptr, dropped);
        } else {  // non-finite or absurdly long segment: the reference's loop would not end; end waypoint only + flag
            step = -1.0;
            bad = true;
        }
    } else if (!(hypot(p0.x - p3.x, p0.y - p3.y) < 1e-1)) {
        bad = true;  // NaN distance
    }
    double *w = ws + g * BEZ_WS;
    w[0] = p1.x; w[1] = p1.y; w[2] = p1.z;
    w[3] = p2.x; w[4] = p2.y; w[5] = p2.z;
    w[6] = step;
    w[7] = (double)rows;
    if (bad && flags) atomicOr(flags + b, BEZ_FLAG_NONFINITE);
}

// Thread per trajectory: start row of every segment (relative to the trajectory) and the trajectory's row count.
__global__ void k_bezier_rows(BatchIdx bi, const double *__restrict__ ws, long long *__restrict__ seg_start,
                              long long *__restrict__ traj_count) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= bi.B) return;
    const long long g0 = bi.seg_begin(b), g1 = bi.seg_begin(b + 1);
    long long total = 0;
    for (long long g = g0; g < g1; ++g) {
        seg_start[g] = total;
        total += (long long)ws[g * BEZ_WS + 7];
    }
    traj_count[b] = total;
}

__global__ void __launch_bounds__(128) k_bezier_write(BatchIdx bi, const double *__restrict__ wp, const double *__restrict__ ws,
                                                      const long long *__restrict__ seg_start,
                                                      const long long *__restrict__ sample_offset, long long capacity,
                                                      double *__restrict__ samples, unsigned *__restrict__ flags) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    const double *p = wp + 3 * (g + b);
    const double *w = ws + g * BEZ_WS;
    const long long row = sample_offset[b] + seg_start[g];
    const double step = w[6];
    bool dropped = false;
    if (step < 0.0) {  // end waypoint only
        if (row < capacity) {
            samples[3 * row] = p[3]; samples[3 * row + 1] = p[4]; samples[3 * row + 2] = p[5];
        } else {
            dropped = true;
        }
    } else {
        const BezPoint p0{p[0], p[1], p[2], 0.0}, p3{p[3], p[4], p[5], 0.0};
        const BezPoint p1{w[0], w[1], w[2], 0.0}, p2{w[3], w[4], w[5], 0.0};
        bez_walk<true>(step, p0, p1, p2, p3, k > 0, row, capacity, samples, dropped);
    }
    if (dropped && flags) atomicOr(flags + b, 2u);  // MSNAP_FLAG_TRUNCATED
}

}  // namespace msnap

#endif  // MSNAP_BEZIER_CUH
