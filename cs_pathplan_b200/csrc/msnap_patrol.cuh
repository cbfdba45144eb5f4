// msnap_patrol.cuh -- post-processing of single-loop patrol trajectories on the sampler's device-resident rows (sm_100a,
// fp64): what UavPathPlanner::gen_single_patrol does with Minisnap_3D's output (SURVEY.md section 8f rank 4;
// /root/reference/uavPathPlanning.cpp:1829-1906 with the helpers at cpp:118-206).
//
// The caller closes the patrol polygon P0..Pn-1 into the waypoint list P0..Pn-1, P0, P1 (cpp:1841-1847) and runs the
// minimum-snap generator on it (cpp:1849).  Per trajectory this stage then
//   1. trims the loop at the sample closest to the second P0: arg-min of the squared 3-D distance over the samples
//      [S/2, S), scanned from the end with a strict `<` (=> the LARGEST index among equal minima)      cpp:1857-1879
//   2. sets every kept sample's `up` to keep_up and appends the first sample again (closed loop)        cpp:1885-1892
//   3. tests the closed loop for a self-intersection in the horizontal plane                            cpp:152-177, 133-150
//   4. if it intersects, replaces it by the polygon boundary sampled every `distance` metres            cpp:1897-1903, 179-206
// Two passes around an exclusive scan of the per-trajectory row counts: k_patrol_count (CTA per trajectory: arg-min
// reduction, all segment pairs of the loop spread over the CTA's threads, fallback row count) and k_patrol_write.
#ifndef MSNAP_PATROL_CUH
#define MSNAP_PATROL_CUH

#include "msnap_generic.cuh"

namespace msnap {

constexpr int PATROL_THREADS = 128;
constexpr unsigned PATROL_FLAG_EMPTY = 1u;      // the generator returned no rows for this loop (cpp:1850-1855): no output rows
constexpr unsigned PATROL_FLAG_FALLBACK = 4u;   // self-intersection after smoothing: boundary sampling was used (cpp:1897-1903)
constexpr double PATROL_EPS = 1e-6;             // cpp:118, 126, 135

struct P2 {
    double e, n;
};
__device__ __forceinline__ double patrol_cross(P2 a, P2 b, P2 c) {  // cross2D, cpp:122-124
    return (b.e - a.e) * (c.n - a.n) - (b.n - a.n) * (c.e - a.e);
}
__device__ __forceinline__ bool patrol_on_segment(P2 a, P2 b, P2 p) {  // onSegment2D, cpp:126-131
    if (fabs(patrol_cross(a, b, p)) > PATROL_EPS) return false;
    return p.e >= fmin(a.e, b.e) - PATROL_EPS && p.e <= fmax(a.e, b.e) + PATROL_EPS && p.n >= fmin(a.n, b.n) - PATROL_EPS &&
           p.n <= fmax(a.n, b.n) + PATROL_EPS;
}
__device__ __forceinline__ bool patrol_segments_intersect(P2 a1, P2 a2, P2 b1, P2 b2) {  // segmentsIntersect2D, cpp:133-150
    const double eps = PATROL_EPS;
    const double c1 = patrol_cross(a1, a2, b1), c2 = patrol_cross(a1, a2, b2);
    const double c3 = patrol_cross(b1, b2, a1), c4 = patrol_cross(b1, b2, a2);
    const bool proper = ((c1 > eps && c2 < -eps) || (c1 < -eps && c2 > eps)) && ((c3 > eps && c4 < -eps) || (c3 < -eps && c4 > eps));
    if (proper) return true;
    if (fabs(c1) <= eps && patrol_on_segment(a1, a2, b1)) return true;
    if (fabs(c2) <= eps && patrol_on_segment(a1, a2, b2)) return true;
    if (fabs(c3) <= eps && patrol_on_segment(b1, b2, a1)) return true;
    if (fabs(c4) <= eps && patrol_on_segment(b1, b2, a2)) return true;
    return false;
}
__device__ __forceinline__ bool patrol_same_xy(double ae, double an, double be, double bn) {  // sameXYPoint, cpp:118-120
    return hypot(ae - be, an - bn) <= PATROL_EPS;
}

// sampleClosedPolygonBoundary (cpp:179-206) of the polygon poly[0..n) ([x, y, z] rows), sequentially as the reference
// does (each point is compared with the previously emitted one).  WRITE: rows go to out with `up` replaced by keep_up.
template <bool WRITE>
__device__ long long patrol_boundary(const double *poly, int n, double spacing, double keep_up, double *out, long long row,
                                     long long capacity, bool &dropped) {
    if (n < 3) return 0;
    const double sp = spacing > 1e-6 ? spacing : 1.0;
    long long cnt = 0;
    double le = 0.0, ln = 0.0, fe = 0.0, fn = 0.0;  // last / first emitted point
    auto emit = [&](double e, double nn) {
        if (WRITE) {
            if (row + cnt < capacity) {
                double *o = out + 3 * (row + cnt);
                o[0] = e; o[1] = nn; o[2] = keep_up;
            } else {
                dropped = true;
            }
        }
        if (cnt == 0) { fe = e; fn = nn; }
        le = e; ln = nn;
        ++cnt;
    };
    for (int i = 0; i < n; ++i) {
        const double *a = poly + 3 * i, *b = poly + 3 * ((i + 1) % n);
        const double dx = b[0] - a[0], dy = b[1] - a[1];
        const double len = hypot(dx, dy);
        double q = ceil(len / sp);
        if (!(q >= 1.0)) q = 1.0;            // std::max(1, (int)ceil(...)); NaN -> 1
        if (q > 2.0e9) q = 2.0e9;
        const int steps = (int)q;
        for (int k = 0; k < steps; ++k) {
            const double t = (double)k / steps;
            const double e = a[0] + t * dx, nn = a[1] + t * dy;
            if (cnt == 0 || !patrol_same_xy(le, ln, e, nn)) emit(e, nn);
        }
    }
    if (cnt > 0 && !patrol_same_xy(fe, fn, le, ln)) emit(fe, fn);
    return cnt;
}

// CTA per trajectory.  best_idx[b] = index of the last kept sample (-1: no rows); count[b] = output rows; flags[b].
__global__ void __launch_bounds__(PATROL_THREADS) k_patrol_count(BatchIdx bi, const double *__restrict__ wp,
                                                                  const long long *__restrict__ sample_offset,
                                                                  const double *__restrict__ samples, long long sample_cap,
                                                                  const double *__restrict__ keep_up, double spacing,
                                                                  long long *__restrict__ best_idx, long long *__restrict__ count,
                                                                  unsigned *__restrict__ flags) {
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    const long long g0 = bi.seg_begin(b);
    const int n_pts = (int)(bi.seg_begin(b + 1) - g0) + 1;
    const double *P = wp + 3 * (g0 + b);  // closed waypoint list: P0..Pn-1, P0, P1
    const long long r0 = sample_offset[b];
    long long S = sample_offset[b + 1] - r0;
    if (r0 + S > sample_cap) S = sample_cap > r0 ? sample_cap - r0 : 0;  // rows that were never written do not exist
    const double *R = samples + 3 * r0;
    __shared__ double s_d[PATROL_THREADS];
    __shared__ long long s_i[PATROL_THREADS];
    if (S <= 0) {
        if (tid == 0) {
            best_idx[b] = -1;
            count[b] = 0;
            flags[b] = PATROL_FLAG_EMPTY;
        }
        return;
    }
    // ---- 1. closest sample to the second P0 (waypoint n_pts - 2), from S/2 on; the largest index wins ties
    long long best = S - 1;
    if (n_pts > 2) {
        const double *T = P + 3 * (n_pts - 2);
        double dmin = 1.7976931348623157e308;  // numeric_limits<double>::max(): a sample at exactly that distance never wins
        long long imin = -1;
        for (long long i = S - 1 - tid; i >= S / 2; i -= PATROL_THREADS) {  // descending per thread, strict <
            const double dx = R[3 * i] - T[0], dy = R[3 * i + 1] - T[1], dz = R[3 * i + 2] - T[2];
            const double d = dx * dx + dy * dy + dz * dz;
            if (d < dmin) { dmin = d; imin = i; }
        }
        s_d[tid] = dmin;
        s_i[tid] = imin;
        __syncthreads();
        for (int o = PATROL_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) {
                const double d2 = s_d[tid + o];
                const long long i2 = s_i[tid + o];
                if (i2 >= 0 && (s_i[tid] < 0 || d2 < s_d[tid] || (d2 == s_d[tid] && i2 > s_i[tid]))) {
                    s_d[tid] = d2;
                    s_i[tid] = i2;
                }
            }
            __syncthreads();
        }
        if (s_i[0] >= 0) best = s_i[0];  // (no finite-distance sample: best_idx stays S - 1, cpp:1859)
        __syncthreads();
    }
    // ---- 3. self-intersection of the closed loop R[0..best] + R[0] (cpp:152-177): the closing point equals the first one,
    // so n = best + 1 distinct points and n closed segments (i, (i+1) % n)
    const long long n = best + 1;
    bool found = false;
    if (n >= 4) {
        // pairs (i, j), j >= i + 2, except (0, n-1): rows i in chunks of 16, the j of a row dealt to the CTA's threads; the
        // vote after every chunk is the (uniform) early exit
        for (long long i0 = 0; i0 < n && !found; i0 += 16) {
            bool hit = false;
            const long long i1 = i0 + 16 < n ? i0 + 16 : n;
            for (long long i = i0; i < i1; ++i) {
                const P2 a1{R[3 * i], R[3 * i + 1]}, a2{R[3 * ((i + 1) % n)], R[3 * ((i + 1) % n) + 1]};
                for (long long j = i + 2 + tid; j < n; j += PATROL_THREADS) {
                    if (i == 0 && j + 1 == n) continue;
                    const P2 b1{R[3 * j], R[3 * j + 1]}, b2{R[3 * ((j + 1) % n)], R[3 * ((j + 1) % n) + 1]};
                    if (patrol_segments_intersect(a1, a2, b1, b2)) hit = true;
                }
            }
            found = __syncthreads_or(hit) != 0;
        }
    }
    if (tid == 0) {
        unsigned f = 0;
        long long c = best + 2;
        if (found) {
            f = PATROL_FLAG_FALLBACK;
            bool dropped = false;
            c = patrol_boundary<false>(P, n_pts - 2, spacing, 0.0, nullptr, 0, 0, dropped);
        }
        best_idx[b] = best;
        count[b] = c;
        flags[b] = f;
    }
}

__global__ void __launch_bounds__(PATROL_THREADS) k_patrol_write(BatchIdx bi, const double *__restrict__ wp,
                                                                  const long long *__restrict__ sample_offset,
                                                                  const double *__restrict__ samples,
                                                                  const double *__restrict__ keep_up, double spacing,
                                                                  const long long *__restrict__ best_idx,
                                                                  const long long *__restrict__ out_offset, long long capacity,
                                                                  double *__restrict__ out, unsigned *__restrict__ flags) {
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    const long long best = best_idx[b];
    if (best < 0) return;
    const long long g0 = bi.seg_begin(b);
    const int n_pts = (int)(bi.seg_begin(b + 1) - g0) + 1;
    const double *P = wp + 3 * (g0 + b);
    const double *R = samples + 3 * sample_offset[b];
    const long long o0 = out_offset[b];
    // keep_up: the last `up` of the trajectory flown before the patrol if the caller has one, else the polygon's first
    // vertex (cpp:1839)
    const double up = keep_up ? keep_up[b] : P[2];
    bool dropped = false;
    if (flags[b] & PATROL_FLAG_FALLBACK) {
        if (tid == 0) patrol_boundary<true>(P, n_pts - 2, spacing, up, out, o0, capacity, dropped);
    } else {
        for (long long i = tid; i <= best + 1; i += PATROL_THREADS) {
            const long long src = i <= best ? i : 0;  // the closing point is the first sample again (cpp:1890-1891)
            if (o0 + i < capacity) {
                double *o = out + 3 * (o0 + i);
                o[0] = R[3 * src]; o[1] = R[3 * src + 1]; o[2] = up;
            } else {
                dropped = true;
            }
        }
    }
    if (dropped) atomicOr(flags + b, 2u);  // MSNAP_FLAG_TRUNCATED
}

}  // namespace msnap

#endif  // MSNAP_PATROL_CUH
