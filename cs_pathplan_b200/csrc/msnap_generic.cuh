// msnap_generic.cuh -- the general-shape kernel set: any order 2..5, any segment count >= 1, mixed lengths (CSR).
//
// One kernel per phase, intermediates in an HBM workspace laid out per segment.  This path is the reference
// implementation *inside* the product (it defines what the fused small-trajectory kernels must reproduce bit for
// bit) and the path long / ragged batches take.  Phases and the reference lines they replace:
//   k_times       ms.cpp:63-72     segment time allocation
//   k_rows        ms.cpp:247-330, 441-469, 474-511 (M, Q, A, V, R)  -> block-tridiagonal rows, O(1) per row
//   k_thomas      ms.cpp:398, 579  (R_PP^-1)  + reweighting loop ms.cpp:76-90 + deviation test ms.cpp:594-624
//   k_search      ms.cpp:408-439   worst-deviation sample per segment
//   k_coeff       ms.cpp:402-404, 584-591, 626-646  (M^-1 d, PolyCoeff packing)
//   k_count/k_traj_count/scan/k_write  ms.cpp:97-161  distance-thresholded sampler, exact CSR output
//   k_stats       ms.cpp:163-195   climb rate / turn radius
#ifndef MSNAP_GENERIC_CUH
#define MSNAP_GENERIC_CUH

#include "msnap_device.cuh"

namespace msnap {

struct BatchIdx {
    long long B;
    long long n_seg;              // total segments
    int ns_uniform;               // > 0: uniform batch, seg_offset unused
    const long long *seg_offset;  // [B+1] device, or nullptr
    const int *seg_traj = nullptr;  // optional [n_seg] trajectory of every segment (k_seg_traj): replaces the search
    __device__ __forceinline__ long long seg_begin(long long b) const {
        return ns_uniform > 0 ? b * ns_uniform : seg_offset[b];
    }
    // trajectory b and local segment k of global segment g
    __device__ __forceinline__ void locate(long long g, long long &b, int &k, int &ns) const {
        if (ns_uniform > 0) {
            b = g / ns_uniform;
            k = (int)(g - b * ns_uniform);
            ns = ns_uniform;
            return;
        }
        if (seg_traj) {
            b = seg_traj[g];
            k = (int)(g - seg_offset[b]);
            ns = (int)(seg_offset[b + 1] - seg_offset[b]);
            return;
        }
        long long lo = 0, hi = B;  // find b with seg_offset[b] <= g < seg_offset[b+1]
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (seg_offset[mid] <= g) lo = mid; else hi = mid;
        }
        b = lo;
        k = (int)(g - seg_offset[lo]);
        ns = (int)(seg_offset[lo + 1] - seg_offset[lo]);
    }
};

// Warp per trajectory: seg_traj[g] = b for every segment of trajectory b (ragged batches; one pass per call).
__global__ void k_seg_traj(long long B, const long long *__restrict__ seg_offset, int *__restrict__ seg_traj) {
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (w >= B) return;
    const long long g1 = seg_offset[w + 1];
    for (long long g = seg_offset[w] + (threadIdx.x & 31); g < g1; g += 32) seg_traj[g] = (int)w;
}

struct SolveParams {
    double pw;           // path_weight (penalty active iff > 0)
    double vw0;          // initial vel_zero_weight
    int max_iter;        // 10 for GenerateTrajectoryMatrix, 0 for a bare SolveQPClosedForm
    const double *vel;   // [B][2][3] or nullptr
    const double *acc;   // [B][2][3] or nullptr
    double bc[12];       // start_vel, end_vel, start_acc, end_acc used when vel/acc are nullptr
};

template <int O>
__device__ __forceinline__ void boundary_of(const SolveParams &sp, long long b, Boundary<O> &bc) {
    const double *v = sp.vel ? sp.vel + 6 * b : sp.bc;
    const double *a = sp.acc ? sp.acc + 6 * b : sp.bc + 6;
    load_boundary<O>(bc, v, a);
}

// Row accessors of the generic path: contiguous rows in the HBM workspace, field stride 1.
template <int NF>
struct GlobalRows {
    double *p;
    static constexpr int FS = 1;
    static constexpr bool ENABLED = true;  // usable as the solution sink of thomas_backward
    static constexpr bool VEC = NF % 2 == 0;  // rows are 16-byte aligned (workspace arrays are 256-byte aligned): row_load
    using Mem = PlainMem;
    __device__ __forceinline__ double *operator()(int j) const { return p + (size_t)j * NF; }
    // The sweeps are one dependent chain per thread, so a row's loads would otherwise pay the full L2/HBM latency:
    // pull row j into L1 while the previous row is being computed.
    __device__ __forceinline__ void prefetch(int j) const {
        const char *q = reinterpret_cast<const char *>(p + (size_t)j * NF);
#pragma unroll
        for (int o = 0; o < NF * 8 + 127; o += 128) {
#ifndef MSNAP_HOST_EMULATION  // (oracle/structured_cpu.cpp compiles these kernels for the host as a CPU baseline)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(q + o));
#else
            __builtin_prefetch(q + o);
#endif
        }
    }
};
struct GlobalPos {  // waypoint positions of one trajectory, [w][3]
    const double *p;
    __device__ __forceinline__ void operator()(int w, double (&out)[3]) const {
        out[0] = p[3 * w];
        out[1] = p[3 * w + 1];
        out[2] = p[3 * w + 2];
    }
};

// ------------------------------------------------------------------------------------------------ k_times
// T_k = max(|P_{k+1} - P_k| / V_avg, min_time) -- plain IEEE mul/add (no FMA contraction) so the allocated
// times, and with them the sampler's candidate grid, are bit-identical to the reference's (ms.cpp:63-72).
__global__ void k_times(BatchIdx bi, const double *__restrict__ wp, double v_avg, double min_time,
                        double *__restrict__ T) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    const double *p = wp + 3 * (g + b);
    const double dx = __dsub_rn(p[3], p[0]), dy = __dsub_rn(p[4], p[1]), dz = __dsub_rn(p[5], p[2]);
    const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    double t = (v_avg > 1e-6) ? __ddiv_rn(len, v_avg) : min_time;
    if (t < min_time) t = min_time;
    T[g] = t;
}

// ------------------------------------------------------------------------------------------------ k_rows
// Thread per segment g = (b, k).  k >= 1 assembles row j = k (between segments k-1 and k) into base[g];
// with the path penalty active every thread also writes its segment's deviation probe segx[g].
template <int O>
__global__ void k_rows(BatchIdx bi, SolveParams sp, const double *__restrict__ wp, const double *__restrict__ T,
                       bool use_pw, const int *__restrict__ s_star, const double *__restrict__ ht,
                       double *__restrict__ base, double *__restrict__ segx) {
    using D = Dim<O>;
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    const double *p = wp + 3 * (g + b);  // waypoint k of trajectory b
    const double Tc = T[g];
    if (use_pw) {  // probe of segment k: h, L(t*), 1/len
        double ip[2 * O], pT[O], h[2 * O];
        time_powers<O>(Tc, ip, pT);
        const int s = s_star[g];
        hermite_at<O>(ht, s, pT, h);
        double x[D::NSEGX];
#pragma unroll
        for (int i = 0; i < 2 * O; ++i) x[i] = h[i];
        const double tau = (double)s * 0.0625;
        double l2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double d = p[3 + a] - p[a];
            x[2 * O + a] = fma(tau, d, p[a]);
            l2 = fma(d, d, l2);
        }
        const double len = sqrt(l2);
        x[2 * O + 3] = len > 1e-6 ? 1.0 / l2 : 0.0;  // 1/len^2: deviations are compared squared
        row_store<GlobalRows<D::NSEGX>, 0, D::NSEGX>(segx + g * D::NSEGX, x);
    }
    if (k == 0) return;
    Boundary<O> bc;
    boundary_of<O>(sp, b, bc);
    double Pm[3], P0[3], Pp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Pm[a] = p[a - 3];
        P0[a] = p[a];
        Pp[a] = p[a + 3];
    }
    double row[D::NBASE];
    assemble_row<O>(T[g - 1], Tc, Pm, P0, Pp, k == 1, k == ns - 1, bc, use_pw, sp.pw, use_pw ? s_star[g - 1] : 0,
                    use_pw ? s_star[g] : 0, ht, row, 1);
    row_store<GlobalRows<D::NBASE>, 0, D::NBASE>(base + g * D::NBASE, row);
}

// ------------------------------------------------------------------------------------------------ k_thomas
// Thread per trajectory: block-tridiagonal Cholesky for the three axes, and -- when eval_dev -- the reweighting
// loop of ms.cpp:76-90: solve, measure max deviation at the recorded t*, double vel_zero_weight while
// max_dev > 0.2 and iter < max_iter.  The final x stays in state[]; per-trajectory results go to the out arrays.
// Row accessors of the generic path: contiguous rows in the HBM workspace, field stride 1.
template <int O>
__global__ void k_thomas(BatchIdx bi, SolveParams sp, const double *__restrict__ wp, double *base,
                         double *state, double *segx, bool eval_dev, bool use_vw,
                         double *__restrict__ max_dev_out, int *__restrict__ iters_out,
                         double *__restrict__ vw_final_out, unsigned *__restrict__ flags) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= bi.B) return;
    const long long g0 = bi.seg_begin(b);
    const int ns = (int)(bi.seg_begin(b + 1) - g0);
    const int n_rows = ns - 1;
    Boundary<O> bc;
    boundary_of<O>(sp, b, bc);
    double d0[NR], dN[NR];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int r = 1; r < O; ++r) {
            d0[(r - 1) * 3 + a] = bc.y0[a][r];
            dN[(r - 1) * 3 + a] = bc.yN[a][r];
        }
    const GlobalRows<D::NBASE> base_at{base + (g0 + 1) * D::NBASE};
    const GlobalRows<D::NSTATE> state_at{state + (g0 + 1) * D::NSTATE};
    const GlobalRows<D::NSEGX> segx_at{segx + g0 * D::NSEGX};
    const GlobalPos pos{wp + 3 * (g0 + b)};

    double vw = use_vw ? sp.vw0 : 0.0;
    int iter = 0;
    double max_dev = 0.0;
    bool ok = true;
    while (true) {
        const double add00 = vw > 0.0 ? 2.0 * vw : 0.0;
        // the last possible iteration (and a bare solve, max_iter == 0) uses the balanced split, like the two-lane
        // chains of the fused kernel; earlier iterations the one-sided split of the speculative lanes (split_row)
        const int m = split_row(n_rows, iter == sp.max_iter || !eval_dev);  // !eval_dev: the loop ends after this solve
        ok = thomas_forward<O>(n_rows, m, add00, base_at, state_at) && ok;
        // x_j overwrites z_j in the state row (z_j is read before it is replaced)
        max_dev = eval_dev ? thomas_backward<O, true, false>(n_rows, m, state_at, state_at, segx_at, pos, d0, dN)
                           : thomas_backward<O, false, false>(n_rows, m, state_at, state_at, segx_at, pos, d0, dN);
        if (max_dev > 0.2 && iter < sp.max_iter) {
            vw = (vw < 1e-6) ? 0.01 : vw * 2.0;
            ++iter;
        } else {
            break;
        }
    }
    if (max_dev_out) max_dev_out[b] = max_dev;
    if (iters_out) iters_out[b] = iter;
    if (vw_final_out) vw_final_out[b] = vw;
    if (flags && (!ok || !(max_dev == max_dev))) atomicOr(flags + b, 1u);
}

// ------------------------------------------------------------------------------------------------ k_thomas_pair
// ONE solve per trajectory (pass 1, the last reweighting iteration, a bare solve) by a PAIR of adjacent lanes:
// balanced split, lane 0 takes the top half and the split row, lane 1 the bottom half (same arithmetic as k_thomas with
// the balanced split, half the chain latency -- these sweeps are pure latency on CSR / long batches).
template <int O>
__global__ void __launch_bounds__(64) k_thomas_pair(BatchIdx bi, SolveParams sp, const double *__restrict__ wp,
                                                    double *base, double *state, double *segx, bool eval_dev,
                                                    double vw, double *__restrict__ max_dev_out,
                                                    unsigned *__restrict__ flags, int *__restrict__ iters_out = nullptr,
                                                    double *__restrict__ vw_final_out = nullptr) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long b = idx >> 1;
    const int side = (int)(idx & 1);
    if (b >= bi.B) return;
    // lanes of this warp that have a trajectory (pairs are never split: 2B is even and warps start at even indices)
    long long valid = 2 * bi.B - (idx - (threadIdx.x & 31));
    if (valid > (long long)blockDim.x) valid = blockDim.x;  // small batches run with fewer than 32 lanes per warp
    const unsigned pair_mask = valid >= 32 ? 0xffffffffu : (1u << (int)valid) - 1u;
    const long long g0 = bi.seg_begin(b);
    const int ns = (int)(bi.seg_begin(b + 1) - g0);
    const int n = ns - 1;
    Boundary<O> bc;
    boundary_of<O>(sp, b, bc);
    double d0[NR], dN[NR];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int r = 1; r < O; ++r) {
            d0[(r - 1) * 3 + a] = bc.y0[a][r];
            dN[(r - 1) * 3 + a] = bc.yN[a][r];
        }
    const GlobalRows<D::NBASE> base_at{base + (g0 + 1) * D::NBASE};
    const GlobalRows<D::NSTATE> state_at{state + (g0 + 1) * D::NSTATE};
    const GlobalRows<D::NSEGX> segx_at{segx + g0 * D::NSEGX};
    const GlobalPos pos{wp + 3 * (g0 + b)};
    const double add00 = vw > 0.0 ? 2.0 * vw : 0.0;
    bool ok = true;
    double m2 = 0.0;
    // segment counts differ between the trajectories of a warp: every lane passes both warp barriers
    const int m = n > 0 ? split_row(n, true) : 0;
    if (n > 0) ok = elim_half<O>(n, side ? n - 1 - m : m, side != 0, add00, base_at, state_at);
    __threadfence_block();
    __syncwarp(pair_mask);
    if (n > 0 && side == 0) ok = elim_middle<O>(n, m, add00, base_at, state_at) && ok;
    __threadfence_block();
    __syncwarp(pair_mask);
    if (n > 0) {
        m2 = eval_dev ? back_half<O, true, false>(n, m, side != 0, state_at, state_at, segx_at, pos, side ? dN : d0, 0.0)
                      : back_half<O, false, false>(n, m, side != 0, state_at, state_at, segx_at, pos, side ? dN : d0, 0.0);
    } else if (eval_dev && side == 0) {
        m2 = single_segment_dev2<O>(segx_at, pos, d0, dN);
    }
    m2 = fmax(m2, __shfl_xor_sync(pair_mask, m2, 1));
    ok = __shfl_xor_sync(pair_mask, ok ? 1 : 0, 1) != 0 && ok;
    if (side == 0) {
        const double md = eval_dev ? sqrt(m2) : 0.0;
        if (max_dev_out) max_dev_out[b] = md;
        if (iters_out) iters_out[b] = 0;  // a solve outside the reweighting loop
        if (vw_final_out) vw_final_out[b] = vw;
        if (flags && (!ok || !(md == md))) atomicOr(flags + b, 1u);
    }
}

// ------------------------------------------------------------------------------------------------ speculative loop
// The reweighting loop of ms.cpp:76-90 for CSR batches, all iterations at once (same idea as the fused kernel): the
// velocity weights are a fixed sequence, so iteration q of trajectory b is an independent solve.
//   k_thomas_spec : thread per (b, q), q = 0..NIT1-1 (all but the last iteration).  Sweep state in a workspace laid out
//                   [row][field][q], so the NIT1 lanes of a trajectory read the same base row (one broadcast load)
//                   and store NIT1 consecutive doubles.  Only the max deviation is kept, and the backward sweep stops at
//                   the first segment safely above the 0.2 threshold (EARLY).
//   k_thomas      : (above) solves the LAST iteration into the final state rows -- where almost every trajectory ends.
//   k_spec_select : thread per trajectory: first iteration whose max_dev <= 0.2, or the last (ms.cpp:82); a trajectory that
//                   stopped early replays that lane's backward sweep into the final state rows.
template <int O, int NIT1>
struct SpecRows {
    double *p;  // this lane's column of the first row
    static constexpr int FS = NIT1;
    static constexpr bool ENABLED = true;
    static constexpr bool VEC = false;
    using Mem = PlainMem;
    __device__ __forceinline__ double *operator()(int j) const { return p + (size_t)j * (Dim<O>::NSTATE * NIT1); }
    __device__ __forceinline__ void prefetch(int) const {}
};

__host__ __device__ inline double reweighted_vw(double vw0, int q) {  // vel_zero_weight of iteration q (ms.cpp:83-87)
    double vw = vw0;
    for (int i = 0; i < q; ++i) vw = (vw < 1e-6) ? 0.01 : vw * 2.0;
    return vw;
}

template <int O, int NIT1>
__global__ void __launch_bounds__(128) k_thomas_spec(BatchIdx bi, SolveParams sp, const double *__restrict__ wp,
                                                     double *base, double *spec_state, double *segx,
                                                     double *__restrict__ md_ws, int *__restrict__ ok_ws) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= bi.B * NIT1) return;
    const long long b = idx / NIT1;
    const int q = (int)(idx - b * NIT1);
    const long long g0 = bi.seg_begin(b);
    const int ns = (int)(bi.seg_begin(b + 1) - g0);
    const int n_rows = ns - 1;
    Boundary<O> bc;
    boundary_of<O>(sp, b, bc);
    double d0[NR], dN[NR];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int r = 1; r < O; ++r) {
            d0[(r - 1) * 3 + a] = bc.y0[a][r];
            dN[(r - 1) * 3 + a] = bc.yN[a][r];
        }
    const GlobalRows<D::NBASE> base_at{base + (g0 + 1) * D::NBASE};
    const SpecRows<O, NIT1> state_at{spec_state + (size_t)(g0 + 1) * D::NSTATE * NIT1 + q};
    const GlobalRows<D::NSEGX> segx_at{segx + g0 * D::NSEGX};
    const GlobalPos pos{wp + 3 * (g0 + b)};
    const double vw = reweighted_vw(sp.vw0, q);
    const double add00 = vw > 0.0 ? 2.0 * vw : 0.0;
    const int m = split_row(n_rows, false);
    const bool ok = thomas_forward<O>(n_rows, m, add00, base_at, state_at);
    const double md = thomas_backward<O, true, true>(n_rows, m, state_at, NoOut{}, segx_at, pos, d0, dN);
    md_ws[idx] = md;
    ok_ws[idx] = ok ? 1 : 0;
}

template <int O, int NIT1>
__global__ void k_spec_select(BatchIdx bi, SolveParams sp, const double *__restrict__ wp, double *spec_state,
                              double *state, double *segx, const double *__restrict__ md_ws,
                              const int *__restrict__ ok_ws, const double *__restrict__ md_last,
                              const unsigned *__restrict__ flag_last, double *__restrict__ max_dev_out, int *__restrict__ iters_out,
                              double *__restrict__ vw_final_out, unsigned *__restrict__ flags) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= bi.B) return;
    int q = 0;
    while (q < NIT1 && md_ws[b * NIT1 + q] > 0.2) ++q;  // q == NIT1: the last iteration (already in `state`)
    double md = md_last[b];
    bool ok = flag_last[b] == 0u;
    if (q < NIT1) {
        md = md_ws[b * NIT1 + q];  // exact: a sweep that passes the test never left early
        ok = ok_ws[b * NIT1 + q] != 0;
        const long long g0 = bi.seg_begin(b);
        const int n_rows = (int)(bi.seg_begin(b + 1) - g0) - 1;
        Boundary<O> bc;
        boundary_of<O>(sp, b, bc);
        double d0[NR], dN[NR];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int r = 1; r < O; ++r) {
                d0[(r - 1) * 3 + a] = bc.y0[a][r];
                dN[(r - 1) * 3 + a] = bc.yN[a][r];
            }
        const SpecRows<O, NIT1> from{spec_state + (size_t)(g0 + 1) * D::NSTATE * NIT1 + q};
        const GlobalRows<D::NSTATE> to{state + (g0 + 1) * D::NSTATE};
        const GlobalRows<D::NSEGX> segx_at{segx + g0 * D::NSEGX};
        const GlobalPos pos{wp + 3 * (g0 + b)};
        thomas_backward<O, false, false>(n_rows, split_row(n_rows, false), from, to, segx_at, pos, d0, dN);
    }
    if (max_dev_out) max_dev_out[b] = md;
    if (iters_out) iters_out[b] = q;
    if (vw_final_out) vw_final_out[b] = reweighted_vw(sp.vw0, q);
    if (flags && (!ok || !(md == md))) atomicOr(flags + b, 1u);  // status of the selected iteration only
}

// Endpoint derivative vectors of segment g = (b, k) from the boundary data and the solved rows.
template <int O>
__device__ __forceinline__ void load_endpoints(const BatchIdx &bi, const SolveParams &sp, const double *wp,
                                               const double *state, long long g, long long b, int k, int ns,
                                               double (&yk)[3][O], double (&yk1)[3][O]) {
    using D = Dim<O>;
    const double *p = wp + 3 * (g + b);
    Boundary<O> bc;
    boundary_of<O>(sp, b, bc);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        yk[a][0] = p[a];
        yk1[a][0] = p[3 + a];
#pragma unroll
        for (int r = 1; r < O; ++r) {
            yk[a][r] = (k == 0) ? bc.y0[a][r] : state[g * D::NSTATE + D::SX + (r - 1) * 3 + a];
            yk1[a][r] = (k == ns - 1) ? bc.yN[a][r] : state[(g + 1) * D::NSTATE + D::SX + (r - 1) * 3 + a];
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_search
// Thread per segment: the 17-sample arg-max of ms.cpp:408-439 on the pass-1 solution (first strict maximum).
template <int O>
__global__ void k_search(BatchIdx bi, SolveParams sp, const double *__restrict__ wp, const double *__restrict__ T,
                         const double *__restrict__ state, int *__restrict__ s_star) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    double yk[3][O], yk1[3][O];
    load_endpoints<O>(bi, sp, wp, state, g, b, k, ns, yk, yk1);
    double ip[2 * O], pT[O];
    time_powers<O>(T[g], ip, pT);
    double dh[3][2 * O];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < O; ++i) {
            dh[a][i] = pT[i] * yk[a][i];
            dh[a][O + i] = pT[i] * yk1[a][i];
        }
    double best = -1.0;
    int best_s = 0;
#pragma unroll
    for (int s = 0; s <= 16; ++s) {
        const double tau = (double)s * 0.0625;
        double d2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double p = 0.0;
#pragma unroll
            for (int i = 0; i < 2 * O; ++i) p = fma(Tab<O>::HT(s, i), dh[a][i], p);
            const double dd = p - fma(tau, yk1[a][0] - yk[a][0], yk[a][0]);
            d2 = fma(dd, dd, d2);
        }
        if (d2 > best) {
            best = d2;
            best_s = s;
        }
    }
    s_star[g] = best_s;
}

// ------------------------------------------------------------------------------------------------ k_coeff
// Thread per segment: polynomial coefficients, PolyCoeff row layout [axis][power hi->lo] (ms.cpp:626-646).
template <int O>
__global__ void k_coeff(BatchIdx bi, SolveParams sp, const double *__restrict__ wp, const double *__restrict__ T,
                        const double *__restrict__ state, double *__restrict__ coeff, unsigned *__restrict__ flags) {
    constexpr int M = 2 * O;
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    double yk[3][O], yk1[3][O];
    load_endpoints<O>(bi, sp, wp, state, g, b, k, ns, yk, yk1);
    double ip[2 * O], pT[O];
    time_powers<O>(T[g], ip, pT);
    bool finite = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double c[M];
        hermite_coeffs<O>(yk[a], yk1[a], ip, pT, c);
        double2 *dst = reinterpret_cast<double2 *>(coeff + (g * 3 + a) * M);
#pragma unroll
        for (int i = 0; i < M; i += 2) {
            dst[i / 2] = make_double2(c[i], c[i + 1]);
            finite = finite && (fabs(c[i]) <= 1.7976931348623157e308) && (fabs(c[i + 1]) <= 1.7976931348623157e308);
        }
    }
    if (flags && !finite) atomicOr(flags + b, 1u);
}

// ------------------------------------------------------------------------------------------------ sampler
// Candidate times of a segment follow ms.cpp:124-141 exactly: dt = min(0.1, T/10), t accumulated by repeated
// addition (NOT i*dt), loop while t <= T + 1e-12, evaluated at min(t, T).
//
// The sampler is split into a COUNT pass and a WRITE pass around an exclusive scan of the per-trajectory row counts.
// The count pass walks the candidates of a segment in order (the acceptance test depends on the last accepted point,
// ms.cpp:143-151) and records WHICH candidates were accepted in a 128-bit mask, so the write pass evaluates only those
// -- at the tabulated times t_i (t_table[i] holds the reference's accumulated `t += 0.1` sequence bit for bit) or, for
// segments shorter than 1 s (dt = T/10), by redoing the <= 11 additions.  Segments with more than 128 candidates
// (T > 12.8 s) fall back to re-running the acceptance loop in the write pass.  Both passes evaluate with the same
// device code, so the rows are identical whichever path wrote them.
__device__ __forceinline__ double sample_dt(double T) {
    double dt = 0.1;
    const double t10 = __ddiv_rn(T, 10.0);
    if (dt > t10) dt = t10;
    return dt;
}

// A segment whose duration is not a positive finite number of at most MSNAP_MAX_SEGMENT_TIME seconds (an inf / NaN /
// absurd waypoint, or min_time_s <= 0 with coincident waypoints: T = 0 => dt = 0) would keep the reference's candidate
// loop `t += dt` spinning forever.  The reference only loses that one call; here one such row would hang the launch of
// the whole batch, so the sampler treats the segment as failed: no candidates, trajectory flagged MSNAP_FLAG_NONFINITE.
constexpr double SAMPLE_T_MAX = 1.0e6;  // == MSNAP_MAX_SEGMENT_TIME (include/msnap.h): at most 1e7 candidates per segment
__device__ __forceinline__ bool sample_time_ok(double T) { return T > 0.0 && T <= SAMPLE_T_MAX; }

constexpr int SAMPLE_MASK_BITS = 128;            // candidates per segment the acceptance mask can describe
constexpr int SAMPLE_TTAB_N = SAMPLE_MASK_BITS + 2;  // entries of t_table a write pass needs
constexpr int SAMPLE_TTAB_BIG = 1 << 14;         // entries of the handle's table in global memory (T up to 1 638 s)

template <int O>
__device__ __forceinline__ void load_coeff(const double *__restrict__ coeff, long long g, double (&c)[3][2 * O]) {
    // coefficient rows are 16*O bytes long and the buffer is 16-byte aligned (checked at the C ABI): 128-bit loads
    const double2 *src = reinterpret_cast<const double2 *>(coeff + g * 3 * 2 * O);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < O; ++i) {
            const double2 v = src[a * O + i];
            c[a][2 * i] = v.x;
            c[a][2 * i + 1] = v.y;
        }
}

// Count pass of one segment.  n: accepted candidates; (m0, m1): acceptance mask of candidates 0..127; usable: every
// candidate has a mask bit; last: the last accepted point (the segment's start point if none was accepted).
// Two candidates are evaluated per trip (their positions do not depend on the acceptance decisions), which doubles the
// instruction-level parallelism of the Horner chains; the decisions are still taken strictly in order.
template <int O>
__device__ __forceinline__ void count_candidates(const double (&c)[3][2 * O], double Tk, const AcceptTest &accept,
                                                 int &n_out, bool &usable, unsigned long long &m0,
                                                 unsigned long long &m1, double (&last)[3]) {
    const double dt = sample_dt(Tk);
    const double tmax = sample_time_ok(Tk) ? Tk + 1e-12 : -1.0;  // a failed segment has no candidates
    double prev[3], ca[3], cb[3];
    eval_xyz<O>(c, 0.0, prev);
    int n = 0, idx = 0;
    unsigned long long w0 = 0ull, w1 = 0ull;
    double t = dt;
    // Dense output (sample_distance <= 0): the test `distance >= 0` holds for every candidate whose position is a number,
    // so the walk only has to count the candidate times.  Taken when no position on [0, T] can overflow or be NaN
    // (|c_i| < 1e100 and T <= 1e6: |p| < 8e142, its square is finite); anything else goes through the tests below.
    if (accept.sd <= 0.0) {
        bool tame = true;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int i = 0; i < 2 * O; ++i) tame = tame && fabs(c[a][i]) < 1e100;
        if (tame) {
            double tl = 0.0;
            while (t <= tmax) {
                tl = t;
                ++idx;
                t += dt;
            }
            if (idx > 0) eval_xyz<O>(c, fmin(tl, Tk), prev);
            n_out = idx;
            usable = idx <= SAMPLE_MASK_BITS;
            m0 = idx >= 64 ? ~0ull : (1ull << idx) - 1ull;
            m1 = idx >= 128 ? ~0ull : (idx > 64 ? (1ull << (idx - 64)) - 1ull : 0ull);
            last[0] = prev[0]; last[1] = prev[1]; last[2] = prev[2];
            return;
        }
    }
    while (t <= tmax) {
        const double t2 = t + dt;
        eval_xyz<O>(c, fmin(t, Tk), ca);
        eval_xyz<O>(c, fmin(t2, Tk), cb);
        if (accept(ca, prev)) {
            prev[0] = ca[0]; prev[1] = ca[1]; prev[2] = ca[2];
            ++n;
            if (idx < 64) w0 |= 1ull << idx; else if (idx < 128) w1 |= 1ull << (idx - 64);
        }
        ++idx;
        if (t2 <= tmax) {
            if (accept(cb, prev)) {
                prev[0] = cb[0]; prev[1] = cb[1]; prev[2] = cb[2];
                ++n;
                if (idx < 64) w0 |= 1ull << idx; else if (idx < 128) w1 |= 1ull << (idx - 64);
            }
            ++idx;
        }
        t = t2 + dt;
    }
    n_out = n;
    usable = idx <= SAMPLE_MASK_BITS;
    m0 = w0;
    m1 = w1;
    last[0] = prev[0]; last[1] = prev[1]; last[2] = prev[2];
}

// Write pass of one segment: put(row, point) for every accepted candidate, rows counted from `row`.
//   ttab: t_table[0 .. SAMPLE_TTAB_N) (shared memory in the callers)
template <int O, class Put>
__device__ __forceinline__ void write_candidates(const double (&c)[3][2 * O], double Tk, bool usable,
                                                 unsigned long long m0, unsigned long long m1,
                                                 const AcceptTest &accept, const double *ttab, long long row,
                                                 const Put &put) {
    const double dt = sample_dt(Tk);
    double cur[3];
    if (usable) {
        const bool tabulated = dt == 0.1;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            unsigned long long m = w == 0 ? m0 : m1;
            while (m) {
                const int bit = __ffsll((long long)m) - 1 + 64 * w;
                m &= m - 1;
                double t;
                if (tabulated) {
                    t = ttab[bit + 1];
                } else {
                    t = dt;
                    for (int i = 0; i < bit; ++i) t += dt;
                }
                eval_xyz<O>(c, fmin(t, Tk), cur);
                put(row++, cur);
            }
        }
    } else {
        double prev[3];
        eval_xyz<O>(c, 0.0, prev);
        const double tmax = sample_time_ok(Tk) ? Tk + 1e-12 : -1.0;
        for (double t = dt; t <= tmax; t += dt) {
            eval_xyz<O>(c, fmin(t, Tk), cur);
            if (accept(cur, prev)) {
                prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
                put(row++, cur);
            }
        }
    }
}

// Warp-cooperative form of the acceptance loop for a segment with MANY candidates (more than the 128 the acceptance
// mask describes: T > 12.8 s -- the reference's shipped mission, 22 km legs at 30 m/s, has 7 000 per segment).  One lane
// walking them costs ~50 dependent instructions per candidate; here the warp evaluates 32 consecutive candidates at once
// and then resolves the acceptances in order: the first lane whose distance to the last accepted point reaches the
// threshold is accepted (ballot + find-first), its point becomes the reference point of the lanes behind it, which are
// tested again, until no lane of the block passes.  Same candidates (t accumulated by repeated addition, every lane
// running the 32 additions of a block), same tests in the same order as ms.cpp:138-152, so the rows are the sequential
// loop's bit for bit.  All 32 lanes call with the same arguments; returns the number of accepted candidates and the last
// accepted point (the segment's start point if none).  WRITE: row `row + k` receives the k-th accepted point.
template <int O, bool WRITE>
__device__ __forceinline__ int warp_sample_long(const double (&c)[3][2 * O], double Tk, const AcceptTest &accept,
                                                const double *__restrict__ t_table, long long row, long long capacity,
                                                double *__restrict__ samples, bool &dropped, double (&last)[3]) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double dt = sample_dt(Tk);
    const double tmax = sample_time_ok(Tk) ? Tk + 1e-12 : -1.0;
    double prev[3];
    eval_xyz<O>(c, 0.0, prev);
    int n = 0;
    double tb = 0.0;  // t of the last candidate of the previous block (t_1 = 0 + dt = dt exactly)
    int base = 0;     // candidates before this block
    bool more = true;
    while (more) {
        double mine;
        if (dt == 0.1 && base + 32 < SAMPLE_TTAB_BIG) {  // t_table[i] = the i-th accumulated time (msnap_create)
            mine = __ldg(t_table + base + lane + 1);
            tb = __shfl_sync(FULL, mine, 31);
        } else {  // beyond the table: every lane runs the block's 32 additions
            double t = tb;
            mine = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                t = __dadd_rn(t, dt);
                if (j == lane) mine = t;
            }
            tb = t;
        }
        base += 32;
        const bool valid = mine <= tmax;  // the candidate times increase: the valid lanes are a prefix of the warp
        more = __all_sync(FULL, valid);
        double cur[3];
        eval_xyz<O>(c, fmin(mine, Tk), cur);
        int start = 0;
        while (true) {
            const bool ok = valid && lane >= start && accept(cur, prev);
            const unsigned m = __ballot_sync(FULL, ok);
            if (!m) break;
            const int f = __ffs(m) - 1;
            prev[0] = __shfl_sync(FULL, cur[0], f);
            prev[1] = __shfl_sync(FULL, cur[1], f);
            prev[2] = __shfl_sync(FULL, cur[2], f);
            if (WRITE && lane == f) {
                const long long r = row + n;
                if (r < capacity) {
                    samples[3 * r] = cur[0]; samples[3 * r + 1] = cur[1]; samples[3 * r + 2] = cur[2];
                } else {
                    dropped = true;
                }
            }
            ++n;
            start = f + 1;
        }
    }
    last[0] = prev[0]; last[1] = prev[1]; last[2] = prev[2];
    return n;
}

// Four blocks per trip: the form k_sample_scan<O, true> uses for batches of long legs (the reference's own mission).  Not
// inlined, and not used by the kernels of ordinary batches: a trip needs ~40 more registers than their thread-per-segment
// paths, and either way of sharing a kernel with them (inlined, or as a call) cost the headline batch 2.6 %.
template <int O, bool WRITE>
__device__ __noinline__ int warp_sample_long4(const double (&c_in)[3][2 * O], double Tk, const AcceptTest &accept_in,
                                             const double *__restrict__ t_table, long long row, long long capacity,
                                             double *__restrict__ samples, bool &dropped, double (&last)[3],
                                             double *rec = nullptr, int rec_cap = 0) {
    // private copies: the arguments live in the caller's stack frame, where a store to `samples` could alias them
    double c[3][2 * O];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 2 * O; ++i) c[a][i] = c_in[a][i];
    const AcceptTest accept = accept_in;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double dt = sample_dt(Tk);
    const double tmax = sample_time_ok(Tk) ? Tk + 1e-12 : -1.0;
    double prev[3];
    eval_xyz<O>(c, 0.0, prev);
    int n = 0;
    double tb = 0.0;  // t of the last candidate of the previous block (t_1 = 0 + dt = dt exactly)
    int base = 0;     // candidates before this block
    bool more = true;
    // acceptances among one block's 32 candidates `cur` (lane k = the block's k-th candidate), in order
    auto resolve = [&](const double (&cur)[3], bool valid, double t_mine) {
        int start = 0;
        while (true) {
            const bool ok = valid && lane >= start && accept(cur, prev);
            const unsigned m = __ballot_sync(FULL, ok);
            if (!m) break;
            const int f = __ffs(m) - 1;
            prev[0] = __shfl_sync(FULL, cur[0], f);
            prev[1] = __shfl_sync(FULL, cur[1], f);
            prev[2] = __shfl_sync(FULL, cur[2], f);
            if (WRITE && lane == f) {
                const long long r = row + n;
                if (r < capacity) {
                    samples[3 * r] = cur[0]; samples[3 * r + 1] = cur[1]; samples[3 * r + 2] = cur[2];
                } else {
                    dropped = true;
                }
            }
            if (rec && lane == f && n < rec_cap) rec[n] = t_mine;
            ++n;
            start = f + 1;
        }
    };
    // Tabulated times (dt = 0.1: t_table[i] = the i-th accumulated time, msnap_create): LONG_U blocks per trip.  A block is a
    // short DEPENDENT chain -- time, Horner, distance, compare, ballot, branch: ~600 cycles for ~40 instructions on a lone
    // warp -- and with the reference's spacings (300 m against 3 m between candidates) most blocks accept nothing, so the
    // chains of LONG_U blocks run side by side and one vote says whether any of them has to be resolved; the times of the next
    // trip are loaded during the current one.  Same candidates, same tests against the same reference point, same order.
    constexpr int LONG_U = 4;
    if (dt == 0.1) {
        double ahead[LONG_U];
#pragma unroll
        for (int j = 0; j < LONG_U; ++j) ahead[j] = __ldg(t_table + 32 * j + lane + 1);  // (the table has >= 32 * LONG_U + 1 entries)
        while (more && base + 32 * LONG_U < SAMPLE_TTAB_BIG) {
            double mine[LONG_U], cur[LONG_U][3];
            bool valid[LONG_U];
#pragma unroll
            for (int j = 0; j < LONG_U; ++j) {
                mine[j] = ahead[j];
                valid[j] = mine[j] <= tmax;  // the candidate times increase: the valid candidates are a prefix of the trip
            }
            if (base + 64 * LONG_U < SAMPLE_TTAB_BIG) {
#pragma unroll
                for (int j = 0; j < LONG_U; ++j) ahead[j] = __ldg(t_table + base + 32 * (LONG_U + j) + lane + 1);
            }
#pragma unroll
            for (int j = 0; j < LONG_U; ++j) eval_xyz<O>(c, fmin(mine[j], Tk), cur[j]);
            more = __all_sync(FULL, valid[LONG_U - 1]);
            // The trip's first acceptance = the lowest candidate that passes; it becomes the reference point of the
            // candidates behind it, which are tested again (all blocks at once), until none passes.  The test is
            // AcceptTest's, with its rare branch (a squared distance within 1e-14 of the threshold, or NaN) taken by the
            // whole warp or not at all, so that the common case is straight-line code for the four blocks.
            int from = 0;  // first candidate of the trip (32 * block + lane) not yet decided
            while (true) {
                bool ok[LONG_U], band = false;
#pragma unroll
                for (int j = 0; j < LONG_U; ++j) {
                    const double dx = cur[j][0] - prev[0], dy = cur[j][1] - prev[1], dz = cur[j][2] - prev[2];
                    const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    ok[j] = d2 >= accept.hi;
                    band = band || !(ok[j] || d2 < accept.lo);
                }
                if (__any_sync(FULL, band)) {
#pragma unroll
                    for (int j = 0; j < LONG_U; ++j) ok[j] = accept(cur[j], prev);
                }
                unsigned m[LONG_U];
#pragma unroll
                for (int j = 0; j < LONG_U; ++j) m[j] = __ballot_sync(FULL, ok[j] && valid[j] && 32 * j + lane >= from);
                int j0 = -1;
#pragma unroll
                for (int j = LONG_U - 1; j >= 0; --j)
                    if (m[j]) j0 = j;
                if (j0 < 0) break;
                unsigned mm = 0u;
                double sx = 0.0, sy = 0.0, sz = 0.0, st = 0.0;
#pragma unroll
                for (int j = 0; j < LONG_U; ++j)
                    if (j == j0) {
                        mm = m[j];
                        sx = cur[j][0], sy = cur[j][1], sz = cur[j][2], st = mine[j];
                    }
                const int f = __ffs(mm) - 1;
                prev[0] = __shfl_sync(FULL, sx, f);
                prev[1] = __shfl_sync(FULL, sy, f);
                prev[2] = __shfl_sync(FULL, sz, f);
                if (WRITE && lane == f) {
                    const long long r = row + n;
                    if (r < capacity) {
                        samples[3 * r] = sx; samples[3 * r + 1] = sy; samples[3 * r + 2] = sz;
                    } else {
                        dropped = true;
                    }
                }
                if (rec && lane == f && n < rec_cap) rec[n] = st;  // the accepted candidate's time, for the write pass
                ++n;
                from = 32 * j0 + f + 1;
            }
            tb = __shfl_sync(FULL, mine[LONG_U - 1], 31);
            base += 32 * LONG_U;
        }
    }
    while (more) {  // beyond the table, or dt = T / 10: every lane runs the block's 32 additions
        double t = tb, mine = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            t = __dadd_rn(t, dt);
            if (j == lane) mine = t;
        }
        tb = t;
        base += 32;
        const bool valid = mine <= tmax;
        more = __all_sync(FULL, valid);
        double cur[3];
        eval_xyz<O>(c, fmin(mine, Tk), cur);
        resolve(cur, valid, mine);
    }
    last[0] = prev[0]; last[1] = prev[1]; last[2] = prev[2];
    return n;
}

// A segment is "long" when it has more candidates than the acceptance mask describes: the 129th tabulated candidate time
// (ttab[129]) still lies within T + 1e-12.  (Segments shorter than 1 s use dt = T/10 and have 10 or 11 candidates.)
__device__ __forceinline__ bool sample_is_long(double Tk, const double *ttab) {
    return sample_time_ok(Tk) && sample_dt(Tk) == 0.1 && ttab[SAMPLE_MASK_BITS + 1] <= Tk + 1e-12;
}

// Stores one sample row unless it lies beyond the caller's capacity.
struct RowSink {
    double *__restrict__ samples;
    long long capacity;
    bool *dropped;
    __device__ __forceinline__ void operator()(long long r, const double (&v)[3]) const {
        if (r < capacity) {
            samples[3 * r] = v[0]; samples[3 * r + 1] = v[1]; samples[3 * r + 2] = v[2];
        } else {
            *dropped = true;
        }
    }
};

// Row sink of the thread-per-segment write kernel.  A thread's rows are consecutive, so an even row is held back until
// its successor arrives and the two leave as three 128-bit stores (48 bytes from a 16-byte aligned address): the kernel
// is bound by the number of store transactions, not by bytes.  flush() must be called after the last row.
struct PairRowSink {
    double *__restrict__ samples;
    long long capacity;
    bool *dropped;
    bool vec;          // the sample buffer is 16-byte aligned
    bool *have;        // a row is held back
    long long *r0;
    double *v0;        // [3]
    __device__ __forceinline__ void scalar(long long r, const double *v) const {
        samples[3 * r] = v[0]; samples[3 * r + 1] = v[1]; samples[3 * r + 2] = v[2];
    }
    __device__ __forceinline__ void flush() const {
        if (*have) scalar(*r0, v0);
        *have = false;
    }
    __device__ __forceinline__ void operator()(long long r, const double (&v)[3]) const {
        if (r >= capacity) {
            *dropped = true;
            return;
        }
        if (*have) {
            if (r == *r0 + 1) {
                double2 *dst = reinterpret_cast<double2 *>(samples + 3 * *r0);
                dst[0] = make_double2(v0[0], v0[1]);
                dst[1] = make_double2(v0[2], v[0]);
                dst[2] = make_double2(v[1], v[2]);
                *have = false;
                return;
            }
            flush();
        }
        if (vec && (r & 1) == 0) {
            *r0 = r;
            v0[0] = v[0]; v0[1] = v[1]; v0[2] = v[2];
            *have = true;
        } else {
            scalar(r, v);
        }
    }
};

// ---- generic (CSR) path: one kernel per pass, intermediates in the HBM workspace ----------------------------------
// k_count: thread per segment.  seg_count[g] = n (usable mask) or -n-1 (write pass must re-run the acceptance loop).
template <int O>
__global__ void __launch_bounds__(128) k_count(BatchIdx bi, const double *__restrict__ coeff,
                                               const double *__restrict__ T, double sample_distance,
                                               int *__restrict__ seg_count, unsigned long long *__restrict__ seg_mask,
                                               double *__restrict__ seg_last, unsigned *__restrict__ flags) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    double c[3][2 * O];
    load_coeff<O>(coeff, g, c);
    const AcceptTest accept(sample_distance);
    int n;
    bool usable;
    unsigned long long m0, m1;
    double last[3];
    count_candidates<O>(c, T[g], accept, n, usable, m0, m1, last);
    if (!sample_time_ok(T[g]) && flags) {
        long long b; int k, ns;
        bi.locate(g, b, k, ns);
        atomicOr(flags + b, 1u);
    }
    seg_count[g] = usable ? n : -n - 1;
    seg_mask[2 * g] = m0;
    seg_mask[2 * g + 1] = m1;
    seg_last[3 * g] = last[0]; seg_last[3 * g + 1] = last[1]; seg_last[3 * g + 2] = last[2];
}

// Thread per trajectory: per-segment start rows (relative to the trajectory), the end-point rule, total count.
template <int O>
__global__ void k_traj_count(BatchIdx bi, const double *__restrict__ coeff, const double *__restrict__ T,
                             const int *__restrict__ seg_count, const double *__restrict__ seg_last,
                             long long *__restrict__ seg_start, int *__restrict__ append_end,
                             long long *__restrict__ traj_count) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= bi.B) return;
    const long long g0 = bi.seg_begin(b), g1 = bi.seg_begin(b + 1);
    long long total = 1;  // first point
    long long last_g = -1;
    for (long long g = g0; g < g1; ++g) {
        seg_start[g] = total;
        const int sc = seg_count[g];
        const int c = sc >= 0 ? sc : -sc - 1;
        total += c;
        if (c > 0) last_g = g;
    }
    double back[3], endp[3], c[3][2 * O];
    if (last_g >= 0) {
        back[0] = seg_last[3 * last_g]; back[1] = seg_last[3 * last_g + 1]; back[2] = seg_last[3 * last_g + 2];
    } else {
        load_coeff<O>(coeff, g0, c);
        eval_xyz<O>(c, 0.0, back);
    }
    load_coeff<O>(coeff, g1 - 1, c);
    eval_xyz<O>(c, T[g1 - 1], endp);
    const int app = dist3(back, endp) > 1e-6 ? 1 : 0;
    append_end[b] = app;
    traj_count[b] = total + app;
}

// k_write: thread per segment.  Rows sample_offset[b] + seg_start[g] + i, plus the trajectory's first point (k == 0)
// and the appended end point (k == ns-1); rows >= capacity are dropped and flagged.
// (4 CTAs per SM: compiled for 5 or 6 the kernel spills and runs 1.44 / 2.03 ms instead of 1.31 ms at cfg5)
template <int O>
__global__ void __launch_bounds__(128, 4) k_write(BatchIdx bi, const double *__restrict__ coeff,
                                               const double *__restrict__ T, double sample_distance,
                                               const double *__restrict__ t_table, const int *__restrict__ seg_count,
                                               const unsigned long long *__restrict__ seg_mask,
                                               const long long *__restrict__ seg_start,
                                               const long long *__restrict__ sample_offset,
                                               const int *__restrict__ append_end, long long capacity,
                                               double *__restrict__ samples, unsigned *__restrict__ flags) {
    __shared__ double ttab[SAMPLE_TTAB_N];
    for (int i = threadIdx.x; i < SAMPLE_TTAB_N; i += blockDim.x) ttab[i] = t_table[i];
    __syncthreads();
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= bi.n_seg) return;
    long long b; int k, ns;
    bi.locate(g, b, k, ns);
    double c[3][2 * O], cur[3];
    load_coeff<O>(coeff, g, c);
    const double Tk = T[g];
    const AcceptTest accept(sample_distance);
    bool dropped = false, have = false;
    long long held_row = 0;
    double held[3];
    const PairRowSink put{samples, capacity, &dropped, (reinterpret_cast<unsigned long long>(samples) & 15ull) == 0,
                          &have, &held_row, held};
    const long long row0 = sample_offset[b];
    if (k == 0) {  // very first point of the trajectory (ms.cpp:132-137)
        eval_xyz<O>(c, 0.0, cur);
        put(row0, cur);
    }
    const int sc = seg_count[g];
    write_candidates<O>(c, Tk, sc >= 0, seg_mask[2 * g], seg_mask[2 * g + 1], accept, ttab, row0 + seg_start[g], put);
    if (k == ns - 1 && append_end[b]) {  // end point appended after the last segment (ms.cpp:157-160)
        eval_xyz<O>(c, Tk, cur);
        put(sample_offset[b + 1] - 1, cur);
    }
    put.flush();
    if (dropped && flags) atomicOr(flags + b, 2u);
}

// Exclusive scan of int64 counts into offsets[n+1], three small kernels (n up to millions; traffic negligible).
constexpr int SCAN_BLOCK = 1024;
__global__ void k_scan_reduce(const long long *__restrict__ in, long long n, long long *__restrict__ partial) {
    __shared__ long long sh[32];
    const long long i = blockIdx.x * (long long)SCAN_BLOCK + threadIdx.x;
    long long v = i < n ? in[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sh[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
// single block: partial[] -> exclusive prefix in place; total into *total_out
__global__ void k_scan_partials(long long *__restrict__ partial, int nblk, long long *__restrict__ total_out) {
    __shared__ long long sh[SCAN_BLOCK];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const long long v = i < nblk ? partial[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
            long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblk) partial[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[SCAN_BLOCK - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}
__global__ void k_scan_apply(const long long *__restrict__ in, long long n, const long long *__restrict__ partial,
                             long long *__restrict__ offsets) {
    __shared__ long long sh[SCAN_BLOCK];
    const long long i = blockIdx.x * (long long)SCAN_BLOCK + threadIdx.x;
    const long long v = i < n ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) offsets[i] = partial[blockIdx.x] + sh[threadIdx.x] - v;
    // offsets[n] is written by k_scan_partials (total_out == offsets + n)
}

// ------------------------------------------------------------------------------------------------ k_sample_scan
// Single-launch sampler for uniform batches: count, exclusive scan across the whole batch and write, in one kernel.
// A CTA takes tiles of `tpt` consecutive trajectories in ticket order (atomic counter), so that a tile's
// predecessors have always started; the cross-tile exclusive prefix uses decoupled look-back on one 64-bit status
// word per tile (2 flag bits + 62 value bits), read SCAN_THREADS predecessors at a time by the whole CTA.
//   status[i]: 0 = nothing yet, (v << 2) | 1 = tile aggregate v, (v << 2) | 2 = inclusive prefix v
//
//   A  count   thread per segment, segments dealt to the lanes in order of decreasing candidate count (a counting
//              sort on the candidate-count estimate), so that the lanes of a warp run loops of similar length
//   B  rows    thread per trajectory: start row of every segment, end-point rule, row count
//   C  scan    tile-local scan, publish the aggregate, EXPAND (below), then look back for the tile's first row
//   D  write   one row per lane, 32 consecutive rows per warp step = 768 contiguous bytes, three coalesced stores
//
// EXPAND turns the acceptance masks into one 32-bit descriptor per output row (segment << 8 | code), written by
// the segment's own lane at the row's tile-local position, so that the write pass needs no search: a lane reads
// its descriptor, takes the candidate time from the table and evaluates.  It runs while the CTA would otherwise only
// wait for its predecessors in the look-back.
constexpr int SCAN_THREADS = 128;
constexpr int SCAN_DESC_CAP = 4096;        // descriptors (rows) expanded per chunk
constexpr unsigned DESC_END = 253u;        // code: the appended end point p(T) of the trajectory's last segment
constexpr unsigned DESC_FIRST = 254u;      // code: the trajectory's first point p(0)
constexpr unsigned DESC_SLOW = 252u;       // code: first row of a segment without a usable mask (lane redoes the segment)
constexpr unsigned DESC_SKIP = 251u;       // code: further rows of such a segment (written by the DESC_SLOW lane)
constexpr unsigned DESC_ACCUM = 1u << 31;  // flag: candidate times are accumulated from dt = T/10 (T < 1 s), not tabulated

// bytes of dynamic shared memory for a tile of `tpt` trajectories of `ns` segments; coef_smem: the tile's
// coefficients are staged in shared memory too (row pitch 6*order + 1 doubles, conflict-free for the count pass)
__host__ __device__ inline size_t scan_smem_bytes(int tpt, int ns, int order, bool coef_smem) {
    const size_t seg_cap = (size_t)tpt * ns;
    return SAMPLE_TTAB_N * sizeof(double) + seg_cap * 2 * sizeof(unsigned long long) + seg_cap * 3 * sizeof(double) +
           seg_cap * sizeof(double) + (coef_smem ? seg_cap * (6 * order + 1) * sizeof(double) : 0) +
           (size_t)(tpt + 1) * sizeof(long long) + seg_cap * 4 * sizeof(int) +
           (size_t)tpt * sizeof(int) + SCAN_DESC_CAP * sizeof(unsigned) + 16;
}

// LONGLEGS: the batch's row capacity says its trajectories have thousands of candidates (scan_tpt: one trajectory per tile);
// the warp-walked segments then take warp_sample_long4.  Same rows either way.
template <int O, bool LONGLEGS = false>
__global__ void __launch_bounds__(SCAN_THREADS, 4) k_sample_scan(
    long long B, int ns, int tpt, long long n_tiles, const double *__restrict__ coeff, const double *__restrict__ T,
    double sample_distance, const double *__restrict__ t_table, unsigned long long *status, unsigned int *ticket,
    long long capacity, long long *__restrict__ sample_offset, double *__restrict__ samples,
    unsigned *__restrict__ flags, int coef_smem, int *sm_ctr, long long *phase_clocks) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int seg_cap = tpt * ns;
    constexpr int CP = 6 * O + 1;  // row pitch of the staged coefficients
    double *ttab = reinterpret_cast<double *>(smem_raw);                                  // [SAMPLE_TTAB_N]
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(ttab + SAMPLE_TTAB_N);  // [seg_cap][2]
    double *last = reinterpret_cast<double *>(mask + 2 * (size_t)seg_cap);                // [seg_cap][3]
    double *segT = last + 3 * (size_t)seg_cap;                                            // [seg_cap] segment times
    double *csm = segT + seg_cap;                                                         // [seg_cap][CP] if coef_smem
    long long *traj_base = reinterpret_cast<long long *>(csm + (coef_smem ? (size_t)seg_cap * CP : 0));  // [tpt + 1]
    int *cnt = reinterpret_cast<int *>(traj_base + tpt + 1);                              // [seg_cap] signed count
    int *seg_start = cnt + seg_cap;                                                       // [seg_cap] row in trajectory
    int *perm = seg_start + seg_cap;                                                      // [seg_cap] longest first
    int *long_list = perm + seg_cap;                                                      // [seg_cap] segments with > 128 candidates
    int *append = long_list + seg_cap;                                                    // [tpt]
    unsigned *desc = reinterpret_cast<unsigned *>(append + tpt);                          // [SCAN_DESC_CAP]
    const AcceptTest accept(sample_distance);
    __shared__ double wstage[(SCAN_THREADS / 32) * 96];  // per warp: 32 rows x 3 doubles staged for coalesced stores
    __shared__ int hist[SAMPLE_MASK_BITS + 1];
    __shared__ long long sh_tile;
    __shared__ long long sh_base;
    __shared__ long long sh_part[SCAN_THREADS / 32];
    __shared__ int sh_first[SCAN_THREADS / 32];
    __shared__ int sh_slot, sh_sched[SCAN_THREADS / 32];
    __shared__ int sh_n_long;
    constexpr int LONG_T_CAP = LONGLEGS ? 2048 : 1;  // accepted candidate times of a tile's warp-walked segments (16 KB)
    __shared__ double long_t[LONG_T_CAP];
    for (int i = tid; i < SAMPLE_TTAB_N; i += SCAN_THREADS) ttab[i] = t_table[i];
    // Which chunk of the ranking a warp takes in phase A: (its scheduler + the CTA's arrival order on this SM) mod 4,
    // a Latin square over (scheduler, co-resident CTA), so that every scheduler of the SM gets one chunk of each rank.
    // The scheduler of a warp is its hardware warp slot mod 4; if the CTA's warps do not sit on four distinct
    // schedulers the plain order (warp w takes chunk w) is used.
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        sh_slot = atomicAdd(sm_ctr + smid, 1);
    }
    if (lane == 0) {
        unsigned wslot;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
        sh_sched[wid] = (int)(wslot & 3u);
    }
    __syncthreads();
    int chunk = wid;
    {
        unsigned seen = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) seen |= 1u << sh_sched[w];
        if (seen == (1u << (SCAN_THREADS / 32)) - 1u) chunk = (sh_sched[wid] + sh_slot) & (SCAN_THREADS / 32 - 1);
    }
    // coefficients of tile-local segment i: from the staged copy once phase A has filled it
    auto staged_coeff = [&](long long g0_, int i, double (&c)[3][2 * O]) {
        if (coef_smem) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < 2 * O; ++j) c[a][j] = csm[i * CP + a * 2 * O + j];
        } else {
            load_coeff<O>(coeff, g0_ + i, c);
        }
    };

    while (true) {
        if (tid == 0) {
            sh_tile = (long long)atomicAdd(ticket, 1u);
            sh_n_long = 0;
        }
        for (int i = tid; i <= SAMPLE_MASK_BITS; i += SCAN_THREADS) hist[i] = 0;
        __syncthreads();
        const long long tile = sh_tile;
        if (tile >= n_tiles) break;
        int stamp = 0;
#define SCAN_STAMP()                                                                   \
    do {                                                                               \
        if (phase_clocks && tid == 0 && tile < 4096 && stamp < 8)                      \
            phase_clocks[(4096 + tile) * 16 + stamp++] = clock64();                         \
    } while (0)
        SCAN_STAMP();
        const long long b0 = tile * tpt;
        const int nt = (int)min((long long)tpt, B - b0);
        const long long g0 = b0 * ns;
        const int nseg = nt * ns;
        // ---- counting sort of the tile's segments by candidate-count estimate, longest first
        for (int i = tid; i < nseg; i += SCAN_THREADS) {
            const double Tk = T[g0 + i];
            segT[i] = Tk;
            const int key = (int)fmax(0.0, fmin(Tk * 10.0, (double)SAMPLE_MASK_BITS));  // ~ candidates when dt = 0.1
            seg_start[i] = key;                          // (scratch until phase B)
            cnt[i] = atomicAdd(&hist[key], 1);           // rank inside the bin (scratch until phase A)
        }
        __syncthreads();
        if (tid < 32) {  // hist[k] := number of segments with a larger key (one warp, 129 bins)
            int run = 0;
            for (int base = SAMPLE_MASK_BITS - (SAMPLE_MASK_BITS % 32); base >= 0; base -= 32) {
                const int k = base + lane;
                const int h = k <= SAMPLE_MASK_BITS ? hist[k] : 0;
                int v = h;  // inclusive suffix sum inside the 32 bins, highest lane first
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_down_sync(0xffffffffu, v, o);
                    if (lane + o < 32) v += u;
                }
                if (k <= SAMPLE_MASK_BITS) hist[k] = run + v - h;
                run += __shfl_sync(0xffffffffu, v, 0);
            }
        }
        __syncthreads();
        for (int i = tid; i < nseg; i += SCAN_THREADS) perm[hist[seg_start[i]] + cnt[i]] = i;
        __syncthreads();
        // ---- A: count accepted candidates per segment, remember which and the last accepted point
        // Warps take the ranking in chunks of 32 (homogeneous loop lengths inside a warp); which chunk a warp takes is
        // `chunk` (see above), so that the long chunks of the CTAs sharing an SM land on different schedulers.
        const int rot_tid = (chunk << 5) | lane;
        for (int r0 = 0; r0 < nseg; r0 += 2 * SCAN_THREADS) {  // snake order over the ranking
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int rk = half == 0 ? r0 + rot_tid : r0 + 2 * SCAN_THREADS - 1 - rot_tid;
                if (rk >= nseg) continue;
                const int i = perm[rk];
                if (sample_is_long(segT[i], ttab)) {  // left to a whole warp (below)
                    long_list[atomicAdd(&sh_n_long, 1)] = i;
                    continue;
                }
                double c[3][2 * O];
                load_coeff<O>(coeff, g0 + i, c);
                if (coef_smem) {
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int j = 0; j < 2 * O; ++j) csm[i * CP + a * 2 * O + j] = c[a][j];
                }
                int n;
                bool usable;
                unsigned long long m0, m1;
                double lp[3];
                count_candidates<O>(c, segT[i], accept, n, usable, m0, m1, lp);
                if (!sample_time_ok(segT[i]) && flags) atomicOr(flags + b0 + i / ns, 1u);
                cnt[i] = usable ? n : -n - 1;
                mask[2 * i] = m0;
                mask[2 * i + 1] = m1;
                last[3 * i] = lp[0]; last[3 * i + 1] = lp[1]; last[3 * i + 2] = lp[2];
            }
        }
        __syncthreads();
        // long segments: a warp each (warp_sample_long), count only
        const int n_long = sh_n_long;
        for (int q = wid; q < n_long; q += SCAN_THREADS / 32) {
            const int i = long_list[q];
            double c[3][2 * O], lp[3];
            load_coeff<O>(coeff, g0 + i, c);
            if (coef_smem && lane == 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int j = 0; j < 2 * O; ++j) csm[i * CP + a * 2 * O + j] = c[a][j];
            }
            bool dropped = false;
            // (long legs: the accepted candidates' times are kept for the write pass, LONG_T_CAP / n_long of them per segment)
            const int rec_cap = LONG_T_CAP / n_long;
            const int n = LONGLEGS ? warp_sample_long4<O, false>(c, segT[i], accept, t_table, 0, 0, nullptr, dropped, lp,
                                                                 long_t + q * rec_cap, rec_cap)
                                   : warp_sample_long<O, false>(c, segT[i], accept, t_table, 0, 0, nullptr, dropped, lp);
            if (lane == 0) {
                cnt[i] = -n - 1;  // negative: no acceptance mask; the rows are written by a warp again (phase D)
                mask[2 * i] = 0ull;
                mask[2 * i + 1] = 0ull;
                last[3 * i] = lp[0]; last[3 * i + 1] = lp[1]; last[3 * i + 2] = lp[2];
            }
        }
        if (n_long > 0) __syncthreads();
        SCAN_STAMP();
        // ---- B: per trajectory: segment start rows, end-point rule (ms.cpp:157-160), row count
        if (tid < nt) {
            int total = 1;  // the first point
            int last_seg = -1;
            for (int k = 0; k < ns; ++k) {
                const int i = tid * ns + k;
                const int n = cnt[i] >= 0 ? cnt[i] : -cnt[i] - 1;
                seg_start[i] = total;
                total += n;
                if (n > 0) last_seg = i;
            }
            double back[3], endp[3], c[3][2 * O];
            if (last_seg >= 0) {
                back[0] = last[3 * last_seg]; back[1] = last[3 * last_seg + 1]; back[2] = last[3 * last_seg + 2];
            } else {
                staged_coeff(g0, tid * ns, c);
                eval_xyz<O>(c, 0.0, back);
            }
            const int il = tid * ns + ns - 1;
            staged_coeff(g0, il, c);
            eval_xyz<O>(c, segT[il], endp);
            const int app = dist3(back, endp) > 1e-6 ? 1 : 0;
            append[tid] = app;
            traj_base[tid + 1] = total + app;  // row count, scanned below
        }
        __syncthreads();
        SCAN_STAMP();
        // ---- C: scan inside the tile and publish the aggregate
        if (tid == 0) {
            long long run = 0;
            traj_base[0] = 0;
            for (int t = 0; t < nt; ++t) {
                const long long c = traj_base[t + 1];
                traj_base[t] = run;
                run += c;
            }
            traj_base[nt] = run;  // tile total
            const unsigned long long word = ((unsigned long long)run << 2) | (tile == 0 ? 2ull : 1ull);
            __threadfence();
            atomicExch(status + tile, word);
            if (tile == 0) sh_base = 0;
        }
        __syncthreads();
        const int total = (int)traj_base[nt];
        // ---- EXPAND rows [c0, c0 + SCAN_DESC_CAP) of the tile into descriptors (first chunk before the look-back)
        auto expand = [&](int c0) {
            // segments in ranking order: the lanes of a warp expand masks of similar population
            for (int rk = tid; rk < nseg; rk += SCAN_THREADS) {
                const int i = perm[rk];
                const int t = i / ns, k = i - t * ns;
                const int tb = (int)traj_base[t] - c0;
                if (k == 0 && tb >= 0 && tb < SCAN_DESC_CAP) desc[tb] = ((unsigned)i << 8) | DESC_FIRST;
                if (k == ns - 1 && append[t]) {
                    const int re = (int)traj_base[t + 1] - 1 - c0;
                    if (re >= 0 && re < SCAN_DESC_CAP) desc[re] = ((unsigned)i << 8) | DESC_END;
                }
                int row = tb + seg_start[i];
                const int sc = cnt[i];
                const int n = sc >= 0 ? sc : -sc - 1;
                if (row >= SCAN_DESC_CAP || row + n <= 0) continue;
                if (sc >= 0) {
                    const unsigned head = ((unsigned)i << 8) | (segT[i] < 1.0 ? DESC_ACCUM : 0u);
                    const uint4 mw = *reinterpret_cast<const uint4 *>(mask + 2 * i);  // 128 acceptance bits, 32 at a time
                    const unsigned words[4] = {mw.x, mw.y, mw.z, mw.w};
                    if (row >= 0 && row + n <= SCAN_DESC_CAP) {  // the whole segment lies inside this chunk
                        unsigned *dst = desc + row;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            unsigned m = words[w];
                            while (m) {
                                *dst++ = head | (unsigned)(__ffs((int)m) - 1 + 32 * w);
                                m &= m - 1;
                            }
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            unsigned m = words[w];
                            while (m) {
                                const int bit = __ffs((int)m) - 1 + 32 * w;
                                m &= m - 1;
                                if (row >= 0 && row < SCAN_DESC_CAP) desc[row] = head | (unsigned)bit;
                                ++row;
                            }
                        }
                    }
                } else {
                    for (int j = 0; j < n; ++j, ++row)  // rows of a long segment: written by a whole warp after the chunk loop
                        if (row >= 0 && row < SCAN_DESC_CAP) desc[row] = ((unsigned)i << 8) | DESC_SKIP;
                }
            }
        };
        expand(0);
        // ---- look back for the exclusive prefix of the tile
        if (tile > 0) {
            long long base = 0;
            long long hi = tile;  // predecessors [hi - SCAN_THREADS, hi) are examined per step, newest first
            bool done = false;
            while (!done) {
                const long long idx = hi - 1 - tid;
                unsigned long long w = 2ull;  // lanes beyond the start of the array count as "prefix 0"
                if (idx >= 0) {
                    const volatile unsigned long long *sp = status + idx;
                    do {
                        w = *sp;  // volatile: re-read from L2 every time
                    } while ((w & 3ull) == 0ull);
                }
                // nearest predecessor (smallest tid) that already has an inclusive prefix
                const unsigned ball = __ballot_sync(0xffffffffu, (w & 3ull) == 2ull);
                if (lane == 0) sh_first[wid] = ball ? (tid & ~31) + __ffs(ball) - 1 : -1;
                __syncthreads();
                int first = -1;
                for (int wi = 0; wi < SCAN_THREADS / 32; ++wi)
                    if (sh_first[wi] >= 0) { first = sh_first[wi]; break; }
                // sum the values of lanes [0, first] (all lanes if no prefix in this window)
                long long v = (first < 0 || tid <= first) ? (long long)(w >> 2) : 0;
                if (idx < 0) v = 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) sh_part[wid] = v;
                __syncthreads();
                for (int wi = 0; wi < SCAN_THREADS / 32; ++wi) base += sh_part[wi];
                done = first >= 0;
                hi -= SCAN_THREADS;
                __syncthreads();
            }
            if (tid == 0) {
                sh_base = base;
                __threadfence();
                atomicExch(status + tile, ((unsigned long long)(base + traj_base[nt]) << 2) | 2ull);
            }
        }
        __syncthreads();
        SCAN_STAMP();
        const long long tile_base = sh_base;
        if (tid < nt) sample_offset[b0 + tid] = tile_base + traj_base[tid];
        if (tile == n_tiles - 1 && tid == 0) sample_offset[B] = tile_base + traj_base[nt];
        // ---- D: write the rows of the tile (same arithmetic as the count pass)
        double *stage = wstage + wid * 96;
        for (int c0 = 0; c0 < total; c0 += SCAN_DESC_CAP) {
            if (c0 > 0) {
                __syncthreads();
                expand(c0);
                __syncthreads();
            }
            const int c1 = min(total, c0 + SCAN_DESC_CAP);
            for (int r0 = c0 + wid * 32; r0 < c1; r0 += SCAN_THREADS) {
                const int r = r0 + lane;
                double cur[3] = {0.0, 0.0, 0.0};
                bool hole = false;  // row written by another lane (DESC_SLOW / DESC_SKIP)
                if (r < c1) {
                    const unsigned d = desc[r - c0];
                    const int i = (int)((d & ~DESC_ACCUM) >> 8);
                    const unsigned code = d & 255u;
                    if (code != DESC_SKIP) {
                        double c[3][2 * O];
                        staged_coeff(g0, i, c);
                        const double Tk = segT[i];
                        {
                            double tt;
                            if (code == DESC_FIRST) tt = 0.0;
                            else if (code == DESC_END) tt = Tk;
                            else if (d & DESC_ACCUM) {
                                const double dt = sample_dt(Tk);
                                tt = dt;
                                for (unsigned n = 0; n < code; ++n) tt += dt;
                                tt = fmin(tt, Tk);
                            } else {
                                tt = fmin(ttab[code + 1], Tk);
                            }
                            eval_xyz<O>(c, tt, cur);
                            if (tile_base + r >= capacity && flags) atomicOr(flags + b0 + i / ns, 2u);
                        }
                    } else {
                        hole = true;
                    }
                }
                stage[3 * lane] = cur[0]; stage[3 * lane + 1] = cur[1]; stage[3 * lane + 2] = cur[2];
                const unsigned holes = __ballot_sync(0xffffffffu, hole);
                __syncwarp();
                const long long e0 = 3 * (tile_base + r0);
                long long room = capacity - (tile_base + r0);  // rows of this step that fit the caller's buffer
                if (room > 32) room = 32;
                const int n_el = 3 * min((int)(room > 0 ? room : 0), c1 - r0);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int e = lane + 32 * j;
                    if (e < n_el && !(holes && ((holes >> (e / 3)) & 1u))) samples[e0 + e] = stage[e];
                }
                __syncwarp();
            }
        }
        // long segments: the same warp-cooperative walk again, this time writing the rows
        for (int q = wid; q < n_long; q += SCAN_THREADS / 32) {
            const int i = long_list[q];
            const int t = i / ns;
            double c[3][2 * O], lp[3];
            staged_coeff(g0, i, c);
            bool dropped = false;
            const int rec_cap = LONG_T_CAP / n_long, n_acc = -cnt[i] - 1;
            if (LONGLEGS && n_acc <= rec_cap) {
                // the count pass recorded which candidates it accepted: evaluate those, 32 at a time, instead of walking the
                // segment's thousands of candidates a second time (same times, same evaluation: the same rows)
                const long long row0 = tile_base + traj_base[t] + seg_start[i];
                const double *rec = long_t + q * rec_cap;
                const double Tk = segT[i];
                for (int k = lane; k < n_acc; k += 32) {
                    double cur[3];
                    eval_xyz<O>(c, fmin(rec[k], Tk), cur);
                    const long long r = row0 + k;
                    if (r < capacity) {
                        samples[3 * r] = cur[0]; samples[3 * r + 1] = cur[1]; samples[3 * r + 2] = cur[2];
                    } else {
                        dropped = true;
                    }
                }
            } else if (LONGLEGS)
                warp_sample_long4<O, true>(c, segT[i], accept, t_table, tile_base + traj_base[t] + seg_start[i], capacity, samples,
                                           dropped, lp);
            else
                warp_sample_long<O, true>(c, segT[i], accept, t_table, tile_base + traj_base[t] + seg_start[i], capacity, samples,
                                          dropped, lp);
            if (__any_sync(0xffffffffu, dropped) && lane == 0 && flags) atomicOr(flags + b0 + t, 2u);
        }
        __syncthreads();  // smem is reused by the next tile
        SCAN_STAMP();
    }
#undef SCAN_STAMP
}

// ------------------------------------------------------------------------------------------------ k_stats
// Warp per trajectory: max |dz|/dxy over consecutive samples and min circumradius over consecutive triples
// (ms.cpp:163-195).  Rows beyond the caller's capacity were never written and are skipped.
__global__ void k_stats(long long B, const long long *__restrict__ sample_offset, const double *__restrict__ samples,
                        long long capacity, double *__restrict__ stats) {
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B) return;
    const long long r0 = sample_offset[w];
    long long r1 = sample_offset[w + 1];
    if (r1 > capacity) r1 = capacity;
    double climb = 0.0, radius = 1.0e12;
    for (long long i = r0 + lane; i + 1 < r1; i += 32) {
        const double *p1 = samples + 3 * i, *p2 = p1 + 3;
        const double dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = fabs(p2[2] - p1[2]);
        const double hd = sqrt(dx * dx + dy * dy);
        if (hd > 1e-6) climb = fmax(climb, dz / hd);
        if (i > r0) {
            const double *p0 = p1 - 3;
            const double ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
            const double cx = p2[0] - p0[0], cy = p2[1] - p0[1], cz = p2[2] - p0[2];
            const double a = sqrt(ax * ax + ay * ay + az * az);
            const double bl = sqrt(dx * dx + dy * dy + (p2[2] - p1[2]) * (p2[2] - p1[2]));
            const double c = sqrt(cx * cx + cy * cy + cz * cz);
            const double ux = ay * cz - az * cy, uy = az * cx - ax * cz, uz = ax * cy - ay * cx;
            const double area = 0.5 * sqrt(ux * ux + uy * uy + uz * uz);
            if (area > 1e-8) radius = fmin(radius, (a * bl * c) / (4.0 * area));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        climb = fmax(climb, __shfl_down_sync(0xffffffffu, climb, o));
        radius = fmin(radius, __shfl_down_sync(0xffffffffu, radius, o));
    }
    if (lane == 0) {
        stats[2 * w] = climb;
        stats[2 * w + 1] = radius;
    }
}

// ------------------------------------------------------------------------------------------------ k_bound
// Upper bound of sample rows: candidates per segment (from T alone) + first and last point per trajectory.
__global__ void k_bound(BatchIdx bi, const double *__restrict__ wp, double v_avg, double min_time,
                        unsigned long long *__restrict__ total) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    unsigned long long n = 0;
    if (g < bi.n_seg) {
        long long b; int k, ns;
        bi.locate(g, b, k, ns);
        const double *p = wp + 3 * (g + b);
        const double dx = __dsub_rn(p[3], p[0]), dy = __dsub_rn(p[4], p[1]), dz = __dsub_rn(p[5], p[2]);
        const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        double t = (v_avg > 1e-6) ? __ddiv_rn(len, v_avg) : min_time;
        if (t < min_time) t = min_time;
        const double dt = sample_dt(t);
        n = !sample_time_ok(t) ? 0ull : (dt == 0.1) ? (unsigned long long)(t / 0.1) + 2ull : 12ull;
        if (k == 0) n += 2ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(total, n);
}

}  // namespace msnap

#endif  // MSNAP_GENERIC_CUH
