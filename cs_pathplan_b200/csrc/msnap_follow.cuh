// msnap_follow.cuh -- follower formation trajectories on the device-resident leader rows (sm_100a, fp64): what
// UavPathPlanner::generateFollowerTrajectories does with the sampled (and altitude-optimised) leader trajectory
// (/root/reference/uavPathPlanning.cpp:3931-4398).
//
// Per leader trajectory of N rows [east, north, up]:
//   heading_t   = atan2 of the central difference of the xy rows (one-sided at both ends; (1, 0) for N = 1)   cpp:3983-4003
//   if N > 5:     heading_t := atan2(sum sin, sum cos) over the raw headings t-10 .. t+10 (clipped)           cpp:4006-4025
//   follower f:   a fixed offset (dx, dy) in the leader's body frame (+x forward, +y left) from the formation
//                 model -- 1 V shape, 2 line abreast, 3 trail columns of max_row, 4 triangle; anything else = 1 --
//                 and the formation distance                                                cpp:4097-4106, 4173-4180, 4258-4273, 4339-4358
//   row (f, t):   east/north = leader_xy_t + R(heading_t) (dx, dy),  up = leader up_t                          cpp:4128-4139
// The rows leave as ENU (k_follow_enu), or -- frame 1 -- as WGS84 {lon, lat, alt} (k_follow_lla: the library's ENU -> WGS84
// statements applied to the row in registers, cpp:4141),
// with the reference's t = 0 rule for models 2-4: {start lon, start lat, leader's first up} (cpp:4188-4195).
// CTA per leader trajectory; sin / cos of the headings are staged in a per-row workspace so that the smoothing window
// reads 21 neighbours instead of recomputing them.
#ifndef MSNAP_FOLLOW_CUH
#define MSNAP_FOLLOW_CUH

namespace msnap {

constexpr int FOLLOW_THREADS = 128;
constexpr int FOLLOW_WINDOW = 10;  // cpp:4009

struct FollowParams {
    int model;        // formation_model (cpp:4053-4071)
    int n_followers;
    int max_row;      // uav_formation_max_row (>= 1)
    double distance;  // formation distance after the lower bound of cpp:4044-4051
};

// body-frame offset of follower idx
__device__ __forceinline__ void follow_offset(const FollowParams &p, int idx, double &dx, double &dy) {
    const double d = p.distance;
    if (p.model == 2) {  // line abreast, cpp:4173-4180
        const int row = idx / 2 + 1, side = (idx % 2 == 0) ? 1 : -1;
        dx = 0.0;
        dy = side * row * d;
    } else if (p.model == 3) {  // trail columns, cpp:4258-4273
        const int col = idx / p.max_row, row_in_col = idx % p.max_row;
        dx = -(row_in_col + 1) * d;
        dy = 0.0;
        if (col > 0) {
            const int side = (col % 2 == 1) ? 1 : -1, level = (col + 1) / 2;
            dy = side * level * d;
        }
    } else if (p.model == 4) {  // triangle, cpp:4339-4358
        const int k = idx + 1;
        int row = 1, prev = 0;
        while (prev + (row + 1) < k) {
            prev += row + 1;
            ++row;
        }
        const int pos = k - prev - 1;
        dx = -row * d;
        const double center = (double)row / 2.0;
        dy = (center - pos) * 2.0 * d;
    } else {  // V shape (model 1 and the default), cpp:4097-4106
        const int row = idx / 2 + 1, side = (idx % 2 == 0) ? 1 : -1;
        dx = -row * d;
        dy = side * row * d;
    }
}

// out: [n_followers * rows][3]; trajectory b owns the block starting at n_followers * row_offset[b], follower-major.
// ws_sin / ws_cos: [rows] scratch.
__global__ void __launch_bounds__(FOLLOW_THREADS) k_follow_enu(FollowParams p, long long B, const long long *__restrict__ row_offset,
                                                               const double *__restrict__ leader, long long n_cap,
                                                               double *__restrict__ ws_sin, double *__restrict__ ws_cos,
                                                               double *__restrict__ out, long long out_cap) {
    const long long b = blockIdx.x;
    const long long r0 = row_offset[b];
    long long N = row_offset[b + 1] - r0;
    if (r0 + N > n_cap) N = n_cap > r0 ? n_cap - r0 : 0;  // rows beyond the caller's buffer do not exist
    if (N <= 0) return;
    const double *L = leader + 3 * r0;
    double *sn = ws_sin + r0, *cs = ws_cos + r0;
    // raw headings (cpp:3983-4003)
    for (long long t = threadIdx.x; t < N; t += FOLLOW_THREADS) {
        double dx, dy;
        if (t == 0) {
            if (N > 1) {
                dx = L[3] - L[0];
                dy = L[4] - L[1];
            } else {
                dx = 1.0;  // cos / sin of the initial heading 0 (cpp:3966-3972, 3990-3991)
                dy = 0.0;
            }
        } else if (t == N - 1) {
            dx = L[3 * (N - 1)] - L[3 * (N - 2)];
            dy = L[3 * (N - 1) + 1] - L[3 * (N - 2) + 1];
        } else {
            dx = L[3 * (t + 1)] - L[3 * (t - 1)];
            dy = L[3 * (t + 1) + 1] - L[3 * (t - 1) + 1];
        }
        double s, c;
        sincos(atan2(dy, dx), &s, &c);
        sn[t] = s;
        cs[t] = c;
    }
    __syncthreads();
    const long long o0 = (long long)p.n_followers * r0;
    for (long long t = threadIdx.x; t < N; t += FOLLOW_THREADS) {
        double s = sn[t], c = cs[t];
        if (N > 5) {  // sliding-window smoothing (cpp:4006-4025), summed in the reference's order
            double sum_sin = 0.0, sum_cos = 0.0;
            for (int k = -FOLLOW_WINDOW; k <= FOLLOW_WINDOW; ++k) {
                const long long idx = t + k;
                if (idx >= 0 && idx < N) {
                    sum_sin += sn[idx];
                    sum_cos += cs[idx];
                }
            }
            sincos(atan2(sum_sin, sum_cos), &s, &c);
        }
        const double le = L[3 * t], ln = L[3 * t + 1], up = L[3 * t + 2];
        for (int f = 0; f < p.n_followers; ++f) {
            double dx, dy;
            follow_offset(p, f, dx, dy);
            const long long row = o0 + (long long)f * N + t;
            if (row < out_cap) {
                double *o = out + 3 * row;
                o[0] = le + (c * dx + -s * dy);  // Rt * rel_body, cpp:4131-4134
                o[1] = ln + (s * dx + c * dy);
                o[2] = up + 0.0;                 // rel_up = 0 (cpp:4108, 4139)
            }
        }
    }
}

// The same rows leaving as WGS84 {lon, lat, alt} (frame 1) in ONE pass: k_follow_enu followed by the in-place ENU -> WGS84
// kernel writes the follower rows, reads them back and writes them again (n_followers x the leader's rows, 72 bytes of
// traffic per row); here the conversion (k_enu_to_wgs84's statements, same bits) is applied to the row while it is in
// registers.  Work items are (follower, row) pairs in output order, so a warp's 32 rows are 768 contiguous bytes and leave
// through the warp's shared-memory tile as three coalesced stores (geo_rows_out).  ws_sin2 / ws_cos2: the smoothed headings.
template <bool TRIG>
__global__ void __launch_bounds__(FOLLOW_THREADS) k_follow_lla(FollowParams p, GeoFrame f, long long B,
                                                               const long long *__restrict__ row_offset,
                                                               const double *__restrict__ leader, long long n_cap,
                                                               double *__restrict__ ws_sin, double *__restrict__ ws_cos,
                                                               double *__restrict__ ws_sin2, double *__restrict__ ws_cos2,
                                                               double *__restrict__ out, long long out_cap) {
    __shared__ double sm_all[FOLLOW_THREADS / 32][96];
    const long long b = blockIdx.x;
    const long long r0 = row_offset[b];
    long long N = row_offset[b + 1] - r0;
    if (r0 + N > n_cap) N = n_cap > r0 ? n_cap - r0 : 0;  // rows beyond the caller's buffer do not exist
    if (N <= 0) return;
    const double *L = leader + 3 * r0;
    double *sn = ws_sin + r0, *cs = ws_cos + r0, *sn2 = ws_sin2 + r0, *cs2 = ws_cos2 + r0;
    for (long long t = threadIdx.x; t < N; t += FOLLOW_THREADS) {  // raw headings (cpp:3983-4003), as in k_follow_enu
        double dx, dy;
        if (t == 0) {
            if (N > 1) {
                dx = L[3] - L[0];
                dy = L[4] - L[1];
            } else {
                dx = 1.0;
                dy = 0.0;
            }
        } else if (t == N - 1) {
            dx = L[3 * (N - 1)] - L[3 * (N - 2)];
            dy = L[3 * (N - 1) + 1] - L[3 * (N - 2) + 1];
        } else {
            dx = L[3 * (t + 1)] - L[3 * (t - 1)];
            dy = L[3 * (t + 1) + 1] - L[3 * (t - 1) + 1];
        }
        double s, c;
        sincos(atan2(dy, dx), &s, &c);
        sn[t] = s;
        cs[t] = c;
    }
    __syncthreads();
    for (long long t = threadIdx.x; t < N; t += FOLLOW_THREADS) {  // sliding-window smoothing (cpp:4006-4025)
        double s = sn[t], c = cs[t];
        if (N > 5) {
            double sum_sin = 0.0, sum_cos = 0.0;
            for (int k = -FOLLOW_WINDOW; k <= FOLLOW_WINDOW; ++k) {
                const long long idx = t + k;
                if (idx >= 0 && idx < N) {
                    sum_sin += sn[idx];
                    sum_cos += cs[idx];
                }
            }
            sincos(atan2(sum_sin, sum_cos), &s, &c);
        }
        sn2[t] = s;
        cs2[t] = c;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *sm = sm_all[w];
    const long long o0 = (long long)p.n_followers * r0, items = (long long)p.n_followers * N;
    const long long o_end = o0 + items < out_cap ? o0 + items : out_cap;  // rows the caller's buffer holds
    for (long long i0 = 32ll * w; i0 < items; i0 += FOLLOW_THREADS) {
        const long long i = i0 + lane;
        double lon = 0.0, lat = 0.0, alt = 0.0;
        if (i < items) {
            // (64-bit division is a subroutine; a trajectory's follower rows fit 32 bits unless it has > 2^31 of them)
            const int fi = items < 0x7fffffffll ? (int)((unsigned)i / (unsigned)N) : (int)(i / N);
            const long long t = i - (long long)fi * N;
            double dx, dy;
            follow_offset(p, fi, dx, dy);
            const double s = sn2[t], c = cs2[t];
            const double e = L[3 * t] + (c * dx + -s * dy);  // Rt * rel_body, cpp:4131-4134
            const double no = L[3 * t + 1] + (s * dx + c * dy);
            const double u = L[3 * t + 2] + 0.0;             // rel_up = 0 (cpp:4108, 4139)
            double X, Y, Z;                                  // k_enu_to_wgs84's statements
            if (TRIG) {
                const double ddx = f.Rinv[0] * e + f.Rinv[1] * no + f.Rinv[2] * u;
                const double ddy = f.Rinv[3] * e + f.Rinv[4] * no + f.Rinv[5] * u;
                const double ddz = f.Rinv[6] * e + f.Rinv[7] * no + f.Rinv[8] * u;
                X = f.ref_ecef[0] + ddx, Y = f.ref_ecef[1] + ddy, Z = f.ref_ecef[2] + ddz;
                geo_ecef_to_wgs84(X, Y, Z, lon, lat, alt);
            } else {
                X = f.ref_ecef[0] + fma(f.Rinv[2], u, fma(f.Rinv[1], no, f.Rinv[0] * e));
                Y = f.ref_ecef[1] + fma(f.Rinv[5], u, fma(f.Rinv[4], no, f.Rinv[3] * e));
                Z = f.ref_ecef[2] + fma(f.Rinv[8], u, fma(f.Rinv[7], no, f.Rinv[6] * e));
                geo_ecef_to_wgs84_fast(X, Y, Z, lon, lat, alt);
            }
        }
        sm[3 * lane] = lon;
        sm[3 * lane + 1] = lat;
        sm[3 * lane + 2] = alt;
        geo_rows_out(out, o0 + i0, o_end, sm, lane);
    }
}

// Models 2-4 in the WGS84 frame: row t = 0 of every follower is {its start lon, its start lat, the leader's first up}
// (cpp:4188-4195, 4280-4287, 4365-4372); thread per (trajectory, follower).
__global__ void k_follow_starts(int n_followers, long long B, const long long *__restrict__ row_offset,
                                const double *__restrict__ leader, long long n_cap, const double *__restrict__ starts,
                                double *__restrict__ out, long long out_cap) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B * n_followers) return;
    const long long b = i / n_followers;
    const int f = (int)(i - b * n_followers);
    const long long r0 = row_offset[b];
    long long N = row_offset[b + 1] - r0;
    if (r0 + N > n_cap) N = n_cap > r0 ? n_cap - r0 : 0;
    if (N <= 0) return;
    const long long row = (long long)n_followers * r0 + (long long)f * N;
    if (row >= out_cap) return;
    out[3 * row] = starts[3 * f];
    out[3 * row + 1] = starts[3 * f + 1];
    out[3 * row + 2] = leader[3 * r0 + 2];
}

}  // namespace msnap

#endif  // MSNAP_FOLLOW_CUH
