// msnap_fused.cuh -- fused, persistent solve kernel for uniform batches (every trajectory has `ns` segments).
//
// One launch replaces k_times .. k_coeff of the generic path.  A CTA of FUSED_THREADS threads owns a tile of `tpc`
// consecutive trajectories whose waypoints, block rows and deviation probes live in shared memory; it walks the
// phases below with __syncthreads() between them and loops over tiles (persistent grid, one wave).
//
//   phase          parallel over            work                                                   reference lines
//   load           elements (coalesced)     waypoints -> smem (axis-major)
//   times          (traj, seg)              T_k                                                      ms.cpp:63-72
//   rows (pass 1)  (traj, row)              D_j, U_j, r_j without penalties                          ms.cpp:247-330, 350
//   thomas pass 1  traj                     block-tridiagonal Cholesky, 3 axes                       ms.cpp:357-405
//   search         (traj, seg)              17-sample arg-max of the deviation -> s*                 ms.cpp:408-439
//   rows (pass 2)  (traj, row) / (traj,seg) rows with pw*h h' and g; deviation probes                ms.cpp:441-469, 511-522
//   thomas spec    (traj, reweight iter)    ALL velocity weights of the reweighting loop at once     ms.cpp:76-90, 474-509, 524-592
//                                           + max deviation at the recorded t*                       ms.cpp:594-624
//   select         traj                     first iteration whose max_dev <= 0.2 (or the 11th)       ms.cpp:82
//   coeff          (traj, seg, axis)        c = M^-1 d, PolyCoeff layout, straight to HBM            ms.cpp:584-591, 626-646
//
// Why speculate on the reweighting loop: its velocity weights are a fixed sequence (vw0, then 0.01 or 2x, ...) and
// only the *decision* to continue is data dependent, so the <= 11 factorisations of a trajectory are independent
// problems.  At the headline size (4 096 x 16) one Thomas chain per trajectory would leave 4 096 threads on a
// 148-SM GPU; 11 chains per trajectory give 45 056, which is what fills the FP64 pipes.  Results are bitwise
// identical to the sequential loop (same arithmetic per iteration, same selection rule).
//
// Per-lane factor/solution state (15 doubles per row at order 4) does not fit in shared memory for that many
// chains; it lives in an L2-resident global scratch slot per CTA, lane-major so every access is a coalesced 256 B
// warp transaction.
#ifndef MSNAP_FUSED_CUH
#define MSNAP_FUSED_CUH

#include "msnap_generic.cuh"

namespace msnap {

constexpr int FUSED_THREADS = 128;

struct FusedParams {
    long long B;
    int ns;
    int tpc;           // trajectories per tile
    int nit;           // lanes per trajectory in the speculative phase: max_iter + 1 if pw > 0 else 1
    int lane_stride;   // lanes per state row in the scratch slot (multiple of 32, >= tpc * nit)
    int traj_stride;   // doubles per trajectory block in shared memory (== 1 mod 16: bank-conflict-free broadcast)
    long long n_tiles;
    const double *wp;
    const double *times_in;  // nullptr => allocate from v_avg / min_time
    double v_avg, min_time;
    SolveParams sp;
    const double *ht;        // HT table of this order in global memory
    double *times_out, *coeff_out, *max_dev_out, *vw_final_out;
    int *iters_out, *best_s_out;
    unsigned *flags;
    double *state_ws;        // [gridDim.x][ns-1][NSTATE][lane_stride]
};

template <int O>
struct FusedSmem {
    // offsets in doubles inside one trajectory block
    int oT, oP, oBase, oSegx, oS, size;
    __host__ __device__ FusedSmem(int ns) {
        using D = Dim<O>;
        const int nr = ns - 1;
        oT = 0;
        oP = oT + ns;
        oBase = oP + 3 * (ns + 1);
        oSegx = oBase + D::NBASE * nr;
        oS = oSegx + D::NSEGX * ns;       // ns ints
        size = oS + (ns + 1) / 2;
        size += (17 - (size % 16)) % 16;  // size == 1 (mod 16)
    }
};

// Row storage accessors handed to thomas_forward / thomas_back_step (pointer to field 0 of row j + field stride).
struct SmemRows {
    const double *ptr;  // field 0 of row 0; rows are consecutive doubles, fields `fs` apart
    int fs;
    __device__ __forceinline__ const double *operator()(int j, int &f) const { f = fs; return ptr + j; }
};
struct SlotRows {
    double *ptr;  // this lane's column of the CTA's scratch slot
    int fs;       // lane stride
    int row;      // doubles per row = NSTATE * lane stride
    __device__ __forceinline__ double *operator()(int j, int &f) const { f = fs; return ptr + (size_t)j * row; }
};

// Per-tile context shared by the phase functions.
template <int O>
struct FusedCtx {
    double *smem;
    double *slot;
    FusedSmem<O> L;
    int ns, nr, nit, tstride, lstride;
    long long b0, g0;
    __device__ FusedCtx(int ns_) : L(ns_) {}
    __device__ __forceinline__ double *block(int t) const { return smem + t * tstride; }
};

// Endpoint derivatives of waypoint w: position from smem, derivatives from the boundary data or a lane's solution.
template <int O>
__device__ __forceinline__ void fused_endpoint(const double *P, int ns, const double *slot_lane, int lane_stride, int w,
                                               const Boundary<O> &bc, double (&y)[3][O]) {
    using D = Dim<O>;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        y[a][0] = P[a * (ns + 1) + w];
#pragma unroll
        for (int r = 1; r < O; ++r) {
            if (w == 0) y[a][r] = bc.y0[a][r];
            else if (w == ns) y[a][r] = bc.yN[a][r];
            else y[a][r] = slot_lane[((size_t)(w - 1) * D::NSTATE + D::ND + (r - 1) * 3 + a) * lane_stride];
        }
    }
}

// The phase bodies are deliberately NOT inlined into the persistent tile loop: inside a loop the compiler hoists the
// constant-table reads (c_tab) into registers, which costs ~100 registers and spills; as straight-line functions the
// table entries fold into the DFMA/DMUL instructions as constant-bank operands.

// rows: item (t, j), j = 1..ns-1
template <int O>
__device__ __noinline__ void fused_row_item(const FusedParams &p, const FusedCtx<O> &c, int t, int j, bool with_pw) {
    const int ns = c.ns;
    double *blk = c.block(t);
    const double *P = blk + c.L.oP;
    const int *ss = reinterpret_cast<const int *>(blk + c.L.oS);
    Boundary<O> bc;
    boundary_of<O>(p.sp, c.b0 + t, bc);
    double Pm[3], P0[3], Pp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Pm[a] = P[a * (ns + 1) + j - 1];
        P0[a] = P[a * (ns + 1) + j];
        Pp[a] = P[a * (ns + 1) + j + 1];
    }
    assemble_row<O>(blk[c.L.oT + j - 1], blk[c.L.oT + j], Pm, P0, Pp, j == 1, j == ns - 1, bc, with_pw, p.sp.pw,
                    with_pw ? ss[j - 1] : 0, with_pw ? ss[j] : 0, p.ht, blk + c.L.oBase + (j - 1), c.nr);
}

// pass-1 Thomas of trajectory t (one lane); returns pivot status
template <int O>
__device__ __noinline__ bool fused_pass1_lane(const FusedCtx<O> &c, int t) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    const SmemRows base_at{c.block(t) + c.L.oBase, c.nr};
    const SlotRows state_at{c.slot + t, c.lstride, D::NSTATE * c.lstride};
    const bool ok = thomas_forward<O>(c.nr, 0.0, base_at, state_at);
    double xn[NR], x[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) xn[i] = 0.0;
    for (int j = c.nr - 1; j >= 0; --j) {
        int bfs, sfs;
        const double *bb = base_at(j, bfs);
        double *st = state_at(j, sfs);
        thomas_back_step<O>(bb, bfs, st, sfs, j + 1 < c.nr, xn, x);
#pragma unroll
        for (int i = 0; i < NR; ++i) xn[i] = x[i];
    }
    return ok;
}

// search: item (t, k) -> index of the worst-deviation sample (first strict maximum, ms.cpp:435)
template <int O>
__device__ __noinline__ int fused_search_item(const FusedParams &p, const FusedCtx<O> &c, int t, int k) {
    const int ns = c.ns;
    const double *blk = c.block(t);
    Boundary<O> bc;
    boundary_of<O>(p.sp, c.b0 + t, bc);
    double yk[3][O], yk1[3][O];
    fused_endpoint<O>(blk + c.L.oP, ns, c.slot + t, c.lstride, k, bc, yk);
    fused_endpoint<O>(blk + c.L.oP, ns, c.slot + t, c.lstride, k + 1, bc, yk1);
    double ip[2 * O], pT[O];
    time_powers<O>(blk[c.L.oT + k], ip, pT);
    double dh[3][2 * O];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < O; ++q) {
            dh[a][q] = pT[q] * yk[a][q];
            dh[a][O + q] = pT[q] * yk1[a][q];
        }
    double best = -1.0;
    int best_s = 0;
#pragma unroll
    for (int s = 0; s <= 16; ++s) {
        const double tau = (double)s * 0.0625;
        double d2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < 2 * O; ++q) v = fma(Tab<O>::HT(s, q), dh[a][q], v);
            const double dd = v - fma(tau, yk1[a][0] - yk[a][0], yk[a][0]);
            d2 = fma(dd, dd, d2);
        }
        if (d2 > best) {
            best = d2;
            best_s = s;
        }
    }
    return best_s;
}

// deviation probe of segment (t, k): h, L(t*), 1/len, field-major in smem
template <int O>
__device__ __noinline__ void fused_probe_item(const FusedParams &p, const FusedCtx<O> &c, int t, int k) {
    const int ns = c.ns;
    double *blk = c.block(t);
    const double *P = blk + c.L.oP;
    double ip[2 * O], pT[O], h[2 * O];
    time_powers<O>(blk[c.L.oT + k], ip, pT);
    const int s = reinterpret_cast<const int *>(blk + c.L.oS)[k];
    hermite_at<O>(p.ht, s, pT, h);
    double *x = blk + c.L.oSegx + k;
#pragma unroll
    for (int q = 0; q < 2 * O; ++q) x[q * ns] = h[q];
    const double tau = (double)s * 0.0625;
    double l2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double d = P[a * (ns + 1) + k + 1] - P[a * (ns + 1) + k];
        x[(2 * O + a) * ns] = fma(tau, d, P[a * (ns + 1) + k]);
        l2 = fma(d, d, l2);
    }
    const double len = sqrt(l2);
    x[(2 * O + 3) * ns] = len > 1e-6 ? 1.0 / len : 0.0;
}

// speculative Thomas: lane (t, q) solves with the q-th velocity weight of the reweighting sequence and measures
// the max deviation at the recorded t* (ms.cpp:594-624).  Returns pivot status.
template <int O>
__device__ __noinline__ bool fused_spec_lane(const FusedParams &p, const FusedCtx<O> &c, int lane, int t, int q,
                                             bool use_pw, double *max_dev_out) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    const int ns = c.ns, nr = c.nr;
    double vw = p.sp.vw0;
    for (int i = 0; i < q; ++i) vw = (vw < 1e-6) ? 0.01 : vw * 2.0;
    const double add00 = vw > 0.0 ? 2.0 * vw : 0.0;
    const double *blk = c.block(t);
    const SmemRows base_at{blk + c.L.oBase, nr};
    const SlotRows state_at{c.slot + lane, c.lstride, D::NSTATE * c.lstride};
    const bool ok = thomas_forward<O>(nr, add00, base_at, state_at);
    Boundary<O> bc;
    boundary_of<O>(p.sp, c.b0 + t, bc);
    const double *P = blk + c.L.oP;
    double xn[NR], x[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) xn[i] = 0.0;
    double yk[3][O], yk1[3][O];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        yk1[a][0] = P[a * (ns + 1) + ns];
#pragma unroll
        for (int r = 1; r < O; ++r) yk1[a][r] = bc.yN[a][r];
    }
    double max_dev = 0.0;
    for (int j = nr - 1; j >= -1; --j) {
        if (j >= 0) {
            int bfs, sfs;
            const double *bb = base_at(j, bfs);
            double *st = state_at(j, sfs);
            thomas_back_step<O>(bb, bfs, st, sfs, j + 1 < nr, xn, x);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                yk[a][0] = P[a * (ns + 1) + j + 1];
#pragma unroll
                for (int r = 1; r < O; ++r) yk[a][r] = x[(r - 1) * 3 + a];
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) xn[i] = x[i];
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                yk[a][0] = P[a * (ns + 1)];
#pragma unroll
                for (int r = 1; r < O; ++r) yk[a][r] = bc.y0[a][r];
            }
        }
        if (use_pw) {
            const double ratio = deviation_ratio<O>(blk + c.L.oSegx + (j + 1), ns, yk, yk1);
            if (ratio > max_dev) max_dev = ratio;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int r = 0; r < O; ++r) yk1[a][r] = yk[a][r];
    }
    *max_dev_out = max_dev;
    return ok;
}

// coefficients of (t, k, axis) from the selected lane's solution, 64-byte rows straight to HBM; returns finiteness
template <int O>
__device__ __noinline__ bool fused_coeff_item(const FusedParams &p, const FusedCtx<O> &c, int t, int k, int a, int lane) {
    using D = Dim<O>;
    constexpr int M = D::M;
    const int ns = c.ns;
    const double *blk = c.block(t);
    const double *P = blk + c.L.oP;
    Boundary<O> bc;
    boundary_of<O>(p.sp, c.b0 + t, bc);
    double yk[O], yk1[O];
    yk[0] = P[a * (ns + 1) + k];
    yk1[0] = P[a * (ns + 1) + k + 1];
#pragma unroll
    for (int d = 1; d < O; ++d) {
        yk[d] = (k == 0) ? bc.y0[a][d]
                         : c.slot[((size_t)(k - 1) * D::NSTATE + D::ND + (d - 1) * 3 + a) * c.lstride + lane];
        yk1[d] = (k == ns - 1) ? bc.yN[a][d]
                               : c.slot[((size_t)k * D::NSTATE + D::ND + (d - 1) * 3 + a) * c.lstride + lane];
    }
    double ip[2 * O], pT[O], co[M];
    time_powers<O>(blk[c.L.oT + k], ip, pT);
    hermite_coeffs<O>(yk, yk1, ip, pT, co);
    double2 *dst = reinterpret_cast<double2 *>(p.coeff_out + ((c.g0 + (long long)t * ns + k) * 3 + a) * M);
    bool finite = true;
#pragma unroll
    for (int q = 0; q < M / 2; ++q) {
        dst[q] = make_double2(co[2 * q], co[2 * q + 1]);
        finite = finite && (fabs(co[2 * q]) <= 1.7976931348623157e308) && (fabs(co[2 * q + 1]) <= 1.7976931348623157e308);
    }
    return finite;
}

template <int O>
__global__ void __launch_bounds__(FUSED_THREADS, 3) k_fused_solve(const __grid_constant__ FusedParams p) {
    using D = Dim<O>;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int ns = p.ns, nr = ns - 1, tpc = p.tpc, nit = p.nit;
    FusedCtx<O> c(ns);
    c.smem = smem;
    c.ns = ns;
    c.nr = nr;
    c.nit = nit;
    c.tstride = p.traj_stride;
    c.lstride = p.lane_stride;
    c.slot = p.state_ws + (size_t)blockIdx.x * nr * D::NSTATE * p.lane_stride;
    double *md = smem + (size_t)tpc * p.traj_stride;     // [tpc][nit] max deviation of every speculative solve
    int *okf = reinterpret_cast<int *>(md + tpc * nit);  // [tpc][nit] pivot status
    int *sel = okf + tpc * nit;                          // [tpc] selected iteration
    int *ok1 = sel + tpc;                                // [tpc] pass-1 pivot status
    const bool use_pw = p.sp.pw > 0.0;

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long b0 = tile * tpc;
        const int nt = (int)min((long long)tpc, p.B - b0);  // trajectories in this tile
        const long long g0 = b0 * ns;                       // first global segment of the tile
        c.b0 = b0;
        c.g0 = g0;

        // ---- load: waypoints of the tile, contiguous in HBM -> axis-major rows in smem
        {
            const double *src = p.wp + 3 * (g0 + b0);
            const int n = nt * (ns + 1) * 3;
            for (int i = tid; i < n; i += FUSED_THREADS) {
                const int t = i / (3 * (ns + 1)), r = i - t * 3 * (ns + 1);
                const int k = r / 3, a = r - 3 * k;
                smem[t * c.tstride + c.L.oP + a * (ns + 1) + k] = src[i];
            }
        }
        __syncthreads();
        // ---- times (plain IEEE mul/add: bit-identical to the reference's allocation, ms.cpp:63-72)
        for (int i = tid; i < nt * ns; i += FUSED_THREADS) {
            const int t = i / ns, k = i - t * ns;
            double *blk = c.block(t);
            double Tk;
            if (p.times_in) {
                Tk = p.times_in[g0 + i];
            } else {
                const double *P = blk + c.L.oP;
                const double dx = __dsub_rn(P[k + 1], P[k]);
                const double dy = __dsub_rn(P[(ns + 1) + k + 1], P[(ns + 1) + k]);
                const double dz = __dsub_rn(P[2 * (ns + 1) + k + 1], P[2 * (ns + 1) + k]);
                const double len =
                    __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                Tk = (p.v_avg > 1e-6) ? __ddiv_rn(len, p.v_avg) : p.min_time;
                if (Tk < p.min_time) Tk = p.min_time;
                if (p.times_out) p.times_out[g0 + i] = Tk;
            }
            blk[c.L.oT + k] = Tk;
            if (!use_pw) {
                reinterpret_cast<int *>(blk + c.L.oS)[k] = 0;
                if (p.best_s_out) p.best_s_out[g0 + i] = 0;
            }
        }
        __syncthreads();

        if (use_pw) {
            // ---- pass 1: snap cost only -> worst-deviation sample per segment
            for (int i = tid; i < nt * nr; i += FUSED_THREADS) fused_row_item<O>(p, c, i / nr, i % nr + 1, false);
            __syncthreads();
            if (tid < nt) ok1[tid] = fused_pass1_lane<O>(c, tid) ? 1 : 0;
            __syncthreads();
            for (int i = tid; i < nt * ns; i += FUSED_THREADS) {
                const int t = i / ns, k = i - t * ns;
                const int s = fused_search_item<O>(p, c, t, k);
                reinterpret_cast<int *>(c.block(t) + c.L.oS)[k] = s;
                if (p.best_s_out) p.best_s_out[g0 + i] = s;
            }
            __syncthreads();
            for (int i = tid; i < nt * ns; i += FUSED_THREADS) fused_probe_item<O>(p, c, i / ns, i % ns);
        }
        // ---- rows of the final system
        for (int i = tid; i < nt * nr; i += FUSED_THREADS) fused_row_item<O>(p, c, i / nr, i % nr + 1, use_pw);
        __syncthreads();
        // ---- speculative Thomas over (trajectory, reweighting iteration)
        if (tid < nt * nit) {
            double mdv;
            const bool ok = fused_spec_lane<O>(p, c, tid, tid / nit, tid % nit, use_pw, &mdv);
            md[tid] = mdv;
            okf[tid] = ok ? 1 : 0;
        }
        __syncthreads();
        // ---- select the iteration the sequential loop would have stopped at (ms.cpp:82)
        if (tid < nt) {
            int q = 0;
            double vw = p.sp.vw0;
            while (md[tid * nit + q] > 0.2 && q < nit - 1) {
                vw = (vw < 1e-6) ? 0.01 : vw * 2.0;
                ++q;
            }
            sel[tid] = q;
            const double mdv = md[tid * nit + q];
            const long long b = b0 + tid;
            if (p.max_dev_out) p.max_dev_out[b] = mdv;
            if (p.iters_out) p.iters_out[b] = q;
            if (p.vw_final_out) p.vw_final_out[b] = vw;
            if (p.flags) {
                const bool bad = !okf[tid * nit + q] || !(mdv == mdv) || (use_pw && !ok1[tid]);
                p.flags[b] = bad ? 1u : 0u;
            }
        }
        __syncthreads();
        // ---- coefficients: items (t, k, axis)
        for (int i = tid; i < nt * ns * 3; i += FUSED_THREADS) {
            const int t = i / (ns * 3), r = i - t * ns * 3;
            const int k = r / 3, a = r - 3 * k;
            const bool finite = fused_coeff_item<O>(p, c, t, k, a, t * nit + sel[t]);
            if (!finite && p.flags) atomicOr(p.flags + b0 + t, 1u);
        }
        __syncthreads();  // the tile's smem and state slot are reused by the next tile
    }
}

}  // namespace msnap

#endif  // MSNAP_FUSED_CUH
