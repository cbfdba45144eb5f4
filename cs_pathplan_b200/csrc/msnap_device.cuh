// msnap_device.cuh -- device-side building blocks of the batched minimum-snap solver (sm_100a, fp64).
//
// Mathematical contract: TrajectoryGeneratorTool::SolveQPClosedForm of the reference
// (/root/reference/math_util/minimum_snap.cpp:227-649), restated per segment in endpoint-derivative ("Hermite")
// space so that the reduced system on the free interior derivatives is SPD block-tridiagonal with
// (order-1) x (order-1) blocks:
//
//   segment k, duration T:  cost = d' S d + g' d,  d = [y_k ; y_{k+1}],  y_j = (p, p', .., p^(o-1)) at waypoint j
//     S = T^(1-2o) D S_hat D  +  pw * h h'  +  vw * (e_1 e_1' + e_{o+1} e_{o+1}'),   D = diag(T^(i mod o))
//     h = D * HT[s*]   (Hermite basis at the worst-deviation sample t* = T s*/16, ms.cpp:408-446)
//     g = -2 pw L(t*) h  (entering the stationarity condition un-halved, exactly as ms.cpp:579 does)
//   row j (interior waypoint j):  U_{j-1}' x_{j-1} + D_j x_j + U_j x_{j+1} = r_j,   x_j = y_j[1..o)
//
// Arithmetic is written with explicit fma() and the library is compiled with --fmad=false, so the sequence of
// roundings is fixed by the source and identical wherever these functions are inlined: the fused kernel and the
// per-phase kernels produce the same bits, and batching or sharding never changes a result.
//
// None of this code is derived from the reference's dense formulation (M, C_T, Q as n_coef^2 matrices and
// 19 dense inverses per solve); it computes the same minimiser in O(ns) work.
#ifndef MSNAP_DEVICE_CUH
#define MSNAP_DEVICE_CUH

#include "msnap_tables.h"

#include <type_traits>


namespace msnap {

template <int O>
struct Dim {
    static constexpr int M = 2 * O;             // coefficients per segment per axis
    static constexpr int B = O - 1;             // block size = free derivatives per interior waypoint
    static constexpr int ND = B * (B + 1) / 2;  // packed symmetric diagonal block
    static constexpr int NU = B * B;            // super-diagonal block U_j (couples x_j with x_{j+1})
    static constexpr int NR = 3 * B;            // right-hand sides, [r][axis]
    static constexpr int F_D = 0, F_U = ND, F_R = ND + NU;
    static constexpr int NBASE = ND + NU + NR;  // per-row constant data (order 4: 24 doubles)
    static constexpr int NSTATE = NR + NU;      // per-row sweep state: z_j (or x_j) and W_j = P_j U_j (order 4: 18 doubles)
    static constexpr int SX = 0, SW = NR;       // field offsets of z/x and of W inside a state row
    static constexpr int NSEGX = M + 4;         // per-segment deviation probe: h[M], L(t*)[3], 1/|P_{k+1}-P_k|
    static constexpr int NP = 2 * O - 1;        // inverse powers of T used by S: T^-1 .. T^-(2o-1)
};

__host__ __device__ constexpr int sym(int r, int q) { return r >= q ? r * (r + 1) / 2 + q : q * (q + 1) / 2 + r; }

// Tables of one order as compile-time constants.  Every use has constant indices after unrolling, so an entry
// becomes a literal operand of the DFMA/DMUL that consumes it -- there is no load for the optimiser to hoist out of
// the persistent loops (a __constant__ array would be hoisted into ~100 live registers).  Lane-divergent lookups
// (HT[s*]) go through a global-memory copy instead (hermite_at).
template <int O>
struct Tab {
    __host__ __device__ static constexpr double S(int i, int j) { return MsnapConstTab<O>::S(i, j); }
    __host__ __device__ static constexpr double H(int k, int i) { return MsnapConstTab<O>::H(k, i); }
    __host__ __device__ static constexpr double HT(int s, int i) { return MsnapConstTab<O>::HT(s, i); }
};

// Boundary conditions of one trajectory: fixed derivatives 1..o-1 at the first and the last waypoint
// (velocity, acceleration given; higher ones zero -- ms.cpp:526-555).
template <int O>
struct Boundary {
    double y0[3][O];  // [axis][r], r = 0 unused
    double yN[3][O];
};

template <int O>
__device__ __forceinline__ void load_boundary(Boundary<O> &bc, const double *vel, const double *acc) {
    // vel/acc: [2][3] (row 0 start, row 1 end) or nullptr for zeros
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int r = 0; r < O; ++r) {
            bc.y0[a][r] = 0.0;
            bc.yN[a][r] = 0.0;
        }
        if (O >= 2 && vel) {
            bc.y0[a][1] = vel[a];
            bc.yN[a][1] = vel[3 + a];
        }
        if (O >= 3 && acc) {
            bc.y0[a][2] = acc[a];
            bc.yN[a][2] = acc[3 + a];
        }
    }
}

// ip[e] = T^-e, e = 1..2o-1 (index 0 unused = 1);  pT[r] = T^r, r = 0..o-1
template <int O>
__device__ __forceinline__ void time_powers(double T, double (&ip)[2 * O], double (&pT)[O]) {
    const double inv = 1.0 / T;
    ip[0] = 1.0;
#pragma unroll
    for (int e = 1; e < 2 * O; ++e) ip[e] = ip[e - 1] * inv;
    pT[0] = 1.0;
#pragma unroll
    for (int r = 1; r < O; ++r) pT[r] = pT[r - 1] * T;
}

// Hermite basis of a segment at tau = s/16 in unscaled derivative space: h[i] = T^(i mod o) * HT[s][i].
// `ht` points at the 17 x MAXM table in GLOBAL memory (lane-divergent s would serialise the constant cache).
template <int O>
__device__ __forceinline__ void hermite_at(const double *__restrict__ ht, int s, const double (&pT)[O],
                                           double (&h)[2 * O]) {
#pragma unroll
    for (int i = 0; i < 2 * O; ++i) h[i] = pT[i % O] * __ldg(ht + s * MSNAP_MAXM + i);
}

// ---------------------------------------------------------------------------------------------------------
// Row assembly: D_j, U_j, r_j of interior waypoint j from its two adjacent segments a = j-1 and c = j.
//   Pm, P0, Pp : waypoints j-1, j, j+1;  first/last: j == 1 / j == ns-1
//   use_pw: add the path penalty (needs s_a, s_c);  the velocity penalty enters later as +2*vw on D[0][0]
//   out: row[F_D..], row[F_U..], row[F_R..] written with stride `fs` between fields
// ---------------------------------------------------------------------------------------------------------
template <int O>
__device__ __forceinline__ void assemble_row_p(const double (&ipa)[2 * O], const double (&pTa)[O],
                                               const double (&ipc)[2 * O], const double (&pTc)[O],
                                               const double (&Pm)[3], const double (&P0)[3], const double (&Pp)[3],
                                               bool first, bool last, const Boundary<O> &bc, bool use_pw, double pw,
                                               int s_a, int s_c, const double *__restrict__ ht, double *row, int fs) {
    using D = Dim<O>;
    constexpr int B = D::B, NP = D::NP;
    double ha[2 * O], hc[2 * O];
    double wa[3], wc[3];  // pw * (h . d_fixed - 2 L(t*)) per axis, for segments a and c
    if (use_pw) {
        hermite_at<O>(ht, s_a, pTa, ha);
        hermite_at<O>(ht, s_c, pTc, hc);
        const double ta = (double)s_a * 0.0625, tc = (double)s_c * 0.0625;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            const double La = fma(ta, P0[x] - Pm[x], Pm[x]);
            const double Lc = fma(tc, Pp[x] - P0[x], P0[x]);
            double sa = fma(ha[O], P0[x], ha[0] * Pm[x]);
            double sc = fma(hc[O], Pp[x], hc[0] * P0[x]);
            if (first) {
#pragma unroll
                for (int q = 1; q < O; ++q) sa = fma(ha[q], bc.y0[x][q], sa);
            }
            if (last) {
#pragma unroll
                for (int q = 1; q < O; ++q) sc = fma(hc[O + q], bc.yN[x][q], sc);
            }
            wa[x] = pw * fma(-2.0, La, sa);
            wc[x] = pw * fma(-2.0, Lc, sc);
        }
    }
#pragma unroll
    for (int r = 1; r <= B; ++r) {
#pragma unroll
        for (int q = 1; q <= r; ++q) {  // packed lower triangle of D_j
            double v = fma(Tab<O>::S(r, q), ipc[NP - r - q], Tab<O>::S(O + r, O + q) * ipa[NP - r - q]);
            if (use_pw) v = fma(pw, fma(hc[r], hc[q], ha[O + r] * ha[O + q]), v);
            row[(D::F_D + sym(r - 1, q - 1)) * fs] = v;
        }
#pragma unroll
        for (int q = 1; q <= B; ++q) {
            double v = Tab<O>::S(r, O + q) * ipc[NP - r - q];
            if (use_pw) v = fma(pw * hc[r], hc[O + q], v);
            row[(D::F_U + (r - 1) * B + (q - 1)) * fs] = v;
        }
        const double ca = Tab<O>::S(O + r, O) * ipa[NP - r];  // S_a[o+r][o] = -S_a[o+r][0]
        const double cc = Tab<O>::S(r, O) * ipc[NP - r];      // S_c[r][o]   = -S_c[r][0]
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double acc = fma(cc, Pp[x] - P0[x], ca * (P0[x] - Pm[x]));
            if (first) {
#pragma unroll
                for (int q = 1; q < O; ++q) acc = fma(Tab<O>::S(O + r, q) * ipa[NP - r - q], bc.y0[x][q], acc);
            }
            if (last) {
#pragma unroll
                for (int q = 1; q < O; ++q) acc = fma(Tab<O>::S(r, O + q) * ipc[NP - r - q], bc.yN[x][q], acc);
            }
            if (use_pw) acc = fma(hc[r], wc[x], fma(ha[O + r], wa[x], acc));
            row[(D::F_R + (r - 1) * 3 + x) * fs] = -acc;
        }
    }
}

// Same, computing the powers of the two segment times on the spot (generic path).
template <int O>
__device__ __forceinline__ void assemble_row(double Ta, double Tc, const double (&Pm)[3], const double (&P0)[3],
                                             const double (&Pp)[3], bool first, bool last, const Boundary<O> &bc,
                                             bool use_pw, double pw, int s_a, int s_c,
                                             const double *__restrict__ ht, double *row, int fs) {
    double ipa[2 * O], ipc[2 * O], pTa[O], pTc[O];
    time_powers<O>(Ta, ipa, pTa);
    time_powers<O>(Tc, ipc, pTc);
    assemble_row_p<O>(ipa, pTa, ipc, pTc, Pm, P0, Pp, first, last, bc, use_pw, pw, s_a, s_c, ht, row, fs);
}

// ---------------------------------------------------------------------------------------------------------
// Small dense kernels on (order-1) x (order-1) SPD blocks, everything in registers.
// ---------------------------------------------------------------------------------------------------------
// Inverse of a packed symmetric positive-definite block, written for LATENCY: the block-tridiagonal sweep is one
// long dependency chain per trajectory, so the pivot path is what the kernel waits on.  B <= 3 use the adjugate
// (all cofactors in parallel, ONE reciprocal); every term of a cofactor carries the same powers of the segment
// times, so the formula is as scale-invariant as a Cholesky factorisation of the same block.  B >= 4 falls back to
// Cholesky.  Returns false if the block is not positive definite (or not finite).
// 1/x for the pivot determinants: hardware seed (rcp.approx.ftz.f64, 2^-23) + two Newton steps, error <= 1 ulp for normal x.
// An IEEE-rounded `1.0 / x` compiles to the same seed and steps PLUS a range check and an out-of-line slow path; that
// subroutine call sits in the row loop of every sweep and costs registers (spills) and ~40 cycles of dependent latency per
// row.  Pivots that are not positive normal numbers are reported through the caller's status flag either way.
__device__ __forceinline__ double pivot_rcp(double x) {
#ifdef MSNAP_HOST_EMULATION  // CPU baseline build of these kernels (oracle/structured_cpu.cpp): an IEEE reciprocal
    return 1.0 / x;
#endif
    double r;
#ifndef MSNAP_HOST_EMULATION
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
    r = 0.0;
#endif
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

template <int B>
__device__ __forceinline__ bool sym_inverse(const double (&a)[B * (B + 1) / 2], double (&p)[B * (B + 1) / 2]) {
    if constexpr (B == 1) {
        p[0] = pivot_rcp(a[0]);
        return a[0] > 0.0;
    } else if constexpr (B == 2) {
        const double det = fma(a[0], a[2], -(a[1] * a[1]));
        const double r = pivot_rcp(det);
        p[0] = a[2] * r;
        p[1] = -a[1] * r;
        p[2] = a[0] * r;
        return a[0] > 0.0 && det > 0.0;
    } else if constexpr (B == 3) {
        // a = [a00, a10, a11, a20, a21, a22]
        const double c00 = fma(a[2], a[5], -(a[4] * a[4]));
        const double c10 = fma(a[4], a[3], -(a[1] * a[5]));
        const double c20 = fma(a[1], a[4], -(a[2] * a[3]));
        const double c11 = fma(a[0], a[5], -(a[3] * a[3]));
        const double c21 = fma(a[1], a[3], -(a[0] * a[4]));
        const double c22 = fma(a[0], a[2], -(a[1] * a[1]));
        const double det = fma(a[3], c20, fma(a[1], c10, a[0] * c00));
        const double r = pivot_rcp(det);
        p[0] = c00 * r;
        p[1] = c10 * r;
        p[2] = c11 * r;
        p[3] = c20 * r;
        p[4] = c21 * r;
        p[5] = c22 * r;
        return a[0] > 0.0 && c22 > 0.0 && det > 0.0;
    } else {
        // Cholesky a = G G', then p = G^-T G^-1
        double g[B * (B + 1) / 2];
        bool ok = true;
#pragma unroll
        for (int i = 0; i < B; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double s = a[sym(i, j)];
#pragma unroll
                for (int k = 0; k < j; ++k) s = fma(-g[sym(i, k)], g[sym(j, k)], s);
                if (i == j) {
                    ok = ok && (s > 0.0);
                    g[sym(i, i)] = rsqrt(s);  // reciprocal diagonal
                } else {
                    g[sym(i, j)] = s * g[sym(j, j)];
                }
            }
        }
        double gi[B][B];  // G^-1 (lower)
#pragma unroll
        for (int c = 0; c < B; ++c) {
#pragma unroll
            for (int i = 0; i < B; ++i) {
                if (i < c) { gi[i][c] = 0.0; continue; }
                double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = c; k < i; ++k) s = fma(-g[sym(i, k)], gi[k][c], s);
                gi[i][c] = s * g[sym(i, i)];
            }
        }
#pragma unroll
        for (int i = 0; i < B; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = i; k < B; ++k) s = fma(gi[k][i], gi[k][j], s);
                p[sym(i, j)] = s;
            }
        return ok;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Block-tridiagonal solve of one trajectory for the three axes at once ("block Thomas"), SPD system
//     U_{j-1}' x_{j-1} + D_j x_j + U_j x_{j+1} = r_j .
// Forward sweep:   D'_j = D_j - U_{j-1}' W_{j-1},   r'_j = r_j - U_{j-1}' z_{j-1},
//                  P_j = D'_j^-1 (registers only),   z_j = P_j r'_j,   W_j = P_j U_j
// Backward sweep:  x_j = z_j - W_j x_{j+1}
//   base(j)  : read-only rows (D, U, r);  state(j): z_j [r][axis] and W_j -- the sweep state is what the
//              speculative lanes stream through L2; the solution is written only where the caller asks for it
//              (XOut).
//   add00    : added to D_j[0][0] of every row (2 * vel_zero_weight: one vw from each adjacent segment).
// Storage is abstracted by accessor types:  `const double* operator()(int j)` = field 0 of row j, and a
// COMPILE-TIME field stride `FS` -- the sweep is issue-bound, so every address must be base + immediate.
// ---------------------------------------------------------------------------------------------------------
// How sweep state is read and written.  PlainMem: ordinary loads/stores (shared memory, or the generic path's HBM
// workspace).  L2KeepMem: global-memory state that is written once and read back once a few microseconds later by
// the same thread -- stores and loads carry an L2 evict_last policy so that the streaming inputs/outputs of the batch
// do not push it out to DRAM in between.
struct PlainMem {
    __device__ static __forceinline__ double ld(const double *p) { return *p; }
    __device__ static __forceinline__ void st(double *p, double v) { *p = v; }
    __device__ static __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
    __device__ static __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
};
#ifndef MSNAP_HOST_EMULATION  // (sm_100a cache-hint accesses: no host counterpart; the CPU baseline uses PlainMem only)
struct L2KeepMem {
    __device__ static __forceinline__ unsigned long long policy() {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        return pol;
    }
    __device__ static __forceinline__ double ld(const double *p) {
        double v;
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy()));
        return v;
    }
    __device__ static __forceinline__ void st(double *p, double v) {
        asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy()) : "memory");
    }
    __device__ static __forceinline__ double2 ld2(const double *p) {
        double2 v;
        asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy()));
        return v;
    }
    __device__ static __forceinline__ void st2(double *p, double a, double b) {
        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(policy()) : "memory");
    }
};

#endif

// Accessors may declare `static constexpr int PAIR = LANES`: the row is then laid out [field / 2][lane][2] -- two consecutive
// fields of one lane sit side by side, so a lane moves them as one 128-bit word and a warp's access is 512 contiguous bytes
// (the speculative lanes' L2 scratch: the sweeps are issue-bound and this halves their state stores and loads).
template <class T, class = void>
struct RowPair { static constexpr int LANES = 0; };
template <class T>
struct RowPair<T, std::void_t<decltype(T::PAIR)>> { static constexpr int LANES = T::PAIR; };

// Consecutive fields F0 .. F0+N-1 of a row.  Accessors whose rows are contiguous and 16-byte aligned (Rows::VEC: the
// per-trajectory rows of the generic path in HBM) move them as 128-bit words: those sweeps touch 32 different cache lines
// per warp access and are bound by the number of load/store transactions.  All other accessors go field by field.
template <class Rows, int F0, int N>
__device__ __forceinline__ void row_load(const double *row, double (&out)[N]) {
    if constexpr (RowPair<Rows>::LANES > 0) {
        constexpr int PL = 2 * RowPair<Rows>::LANES, H = F0 & 1;
        if (H) out[0] = Rows::Mem::ld(row + (F0 >> 1) * PL + 1);
#pragma unroll
        for (int k = H; k + 1 < N; k += 2) {
            const double2 v = Rows::Mem::ld2(row + ((F0 + k) >> 1) * PL);
            out[k] = v.x;
            out[k + 1] = v.y;
        }
        if ((N - H) & 1) out[N - 1] = Rows::Mem::ld(row + ((F0 + N - 1) >> 1) * PL);
    } else if constexpr (Rows::VEC) {
        constexpr int H = F0 & 1;
        if (H) out[0] = row[F0];
#pragma unroll
        for (int k = H; k + 1 < N; k += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(row + F0 + k);
            out[k] = v.x;
            out[k + 1] = v.y;
        }
        if ((N - H) & 1) out[N - 1] = row[F0 + N - 1];
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = Rows::Mem::ld(row + (F0 + k) * Rows::FS);
    }
}
template <class Rows, int F0, int N>
__device__ __forceinline__ void row_store(double *row, const double (&in)[N]) {
    if constexpr (RowPair<Rows>::LANES > 0) {
        constexpr int PL = 2 * RowPair<Rows>::LANES, H = F0 & 1;
        if (H) Rows::Mem::st(row + (F0 >> 1) * PL + 1, in[0]);
#pragma unroll
        for (int k = H; k + 1 < N; k += 2) Rows::Mem::st2(row + ((F0 + k) >> 1) * PL, in[k], in[k + 1]);
        if ((N - H) & 1) Rows::Mem::st(row + ((F0 + N - 1) >> 1) * PL, in[N - 1]);
    } else if constexpr (Rows::VEC) {
        constexpr int H = F0 & 1;
        if (H) row[F0] = in[0];
#pragma unroll
        for (int k = H; k + 1 < N; k += 2) *reinterpret_cast<double2 *>(row + F0 + k) = make_double2(in[k], in[k + 1]);
        if ((N - H) & 1) row[F0 + N - 1] = in[N - 1];
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) Rows::Mem::st(row + (F0 + k) * Rows::FS, in[k]);
    }
}

// ---- twisted (two-sided) elimination ---------------------------------------------------------------------------
// The chain of n block rows is split at a row m and eliminated from BOTH ends towards it:
//   top half     rows 0 .. m-1      D'_j = D_j - U_{j-1}' W_{j-1},  z_j = P_j (r_j - U_{j-1}' z_{j-1}),  W_j = P_j U_j
//   bottom half  rows n-1 .. m+1    D'_j = D_j - U_j W_{j+1},       z_j = P_j (r_j - U_j z_{j+1}),       W_j = P_j U_{j-1}'
//   split row m                     D*  = D_m - U_{m-1}' W_{m-1} - U_m W_{m+1},  x_m = D*^-1 (r_m - U_{m-1}' z_{m-1} - U_m z_{m+1})
// and substituted back outwards:  x_j = z_j - W_j x_{j+1} (j < m),  x_j = z_j - W_j x_{j-1} (j > m).
// Every pivot block is a Schur complement of the SPD matrix, hence SPD.  The bottom half is the top-half recurrence
// applied to the MIRRORED chain (row i of the mirrored chain = row n-1-i, coupling blocks transposed), so one routine
// with a `mirror` flag that only enters address computations serves both.  The two halves are independent until the
// split row: two lanes can each take one (the latency-critical chains of the fused kernel do, with m = n/2), or one lane
// runs them in turn -- the operations, and therefore the bits, are the same.
//   m = n/2 : the balanced split (pass 1, a bare solve, the last reweighting iteration)
//   m = n-1 : everything in the top half = plain downward block Thomas; the back-substitution then starts at the LAST
//             segment, whose deviation usually decides the reweighting test at once (speculative iterations).
__host__ __device__ constexpr int split_row(int n_rows, bool balanced) { return balanced ? n_rows / 2 : n_rows - 1; }

// coupling block of a base row in the orientation the recurrences use: C[t][p] = U[t][p], or U[p][t] when transposed
template <int O, class BaseAt>
__device__ __forceinline__ void load_coupling(const double *b, bool transposed, double (&C)[(O - 1) * (O - 1)]) {
    using D = Dim<O>;
    constexpr int B = D::B;
    if constexpr (BaseAt::VEC) {
        double U[B * B];
        row_load<BaseAt, D::F_U, B * B>(b, U);
#pragma unroll
        for (int t = 0; t < B; ++t)
#pragma unroll
            for (int p = 0; p < B; ++p) C[t * B + p] = transposed ? U[p * B + t] : U[t * B + p];
    } else {
#pragma unroll
        for (int t = 0; t < B; ++t)
#pragma unroll
            for (int p = 0; p < B; ++p)
                C[t * B + p] = b[(D::F_U + (transposed ? p * B + t : t * B + p)) * BaseAt::FS];
    }
}

// Elimination of one half: local rows i = 0 .. cnt-1, outermost first; row index j = mirror ? n_rows-1-i : i.
// Returns false if a pivot block was not positive definite.
template <int O, class BaseAt, class StateAt>
__device__ __forceinline__ bool elim_half(int n_rows, int cnt, bool mirror, double add00, const BaseAt base_at,
                                          const StateAt state_at) {
    using D = Dim<O>;
    constexpr int B = D::B, ND = D::ND, NR = D::NR, NU = D::NU;
    bool ok = true;
    double W[NU];  // W of the previously eliminated row
    double z[NR];  // z of the previously eliminated row, [r][axis]
    const int step = mirror ? -1 : 1;
    int j = mirror ? n_rows - 1 : 0;
    // One row; FIRST (a compile-time flag) = the outermost row, which has no previously eliminated neighbour.  The first
    // row is peeled off the loop so that W and z are plain loop-carried values, defined on every path into the loop body.
    auto row = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        base_at.prefetch(j + step);  // the next row towards the split row (which always exists)
        const double *b = base_at(j);
        double *s = state_at(j);
        double d[ND], r[NR];
        row_load<BaseAt, D::F_D, ND>(b, d);
        row_load<BaseAt, D::F_R, NR>(b, r);
        d[0] += add00;
        if constexpr (!FIRST) {
            // coupling to the previously eliminated row: U_{j-1} (stored in row j-1), or U_j' (row j) when mirrored;
            // re-read (a broadcast LDS) rather than carried in registers
            double C[NU];
            load_coupling<O, BaseAt>(mirror ? b : base_at(j - 1), mirror, C);
#pragma unroll
            for (int p = 0; p < B; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) {
                    double acc = d[sym(p, q)];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], W[t * B + q], acc);
                    d[sym(p, q)] = acc;
                }
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double acc = r[p * 3 + x];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], z[t * 3 + x], acc);
                    r[p * 3 + x] = acc;
                }
            }
        }
        double P[ND];
        ok = sym_inverse<B>(d, P) && ok;
#pragma unroll
        for (int p = 0; p < B; ++p)
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double acc = P[sym(p, 0)] * r[x];
#pragma unroll
                for (int t = 1; t < B; ++t) acc = fma(P[sym(p, t)], r[t * 3 + x], acc);
                z[p * 3 + x] = acc;
            }
        row_store<StateAt, D::SX, NR>(s, z);
        {
            // coupling to the next row towards the split row: U_j (row j), or U_{j-1}' (row j-1) when mirrored
            double N[NU];
            load_coupling<O, BaseAt>(mirror ? base_at(j - 1) : b, mirror, N);
#pragma unroll
            for (int p = 0; p < B; ++p)
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    double acc = P[sym(p, 0)] * N[q];
#pragma unroll
                    for (int t = 1; t < B; ++t) acc = fma(P[sym(p, t)], N[t * B + q], acc);
                    W[p * B + q] = acc;
                }
            row_store<StateAt, D::SW, NU>(s, W);
        }
        j += step;
    };
    if (cnt <= 0) return ok;
    row(std::true_type{});
    for (int i = 1; i < cnt; ++i) row(std::false_type{});
    return ok;
}

// The split row m receives both halves (from the state rows m-1 and m+1) and is solved for x_m, which is left in the SX
// field of state row m.
template <int O, class BaseAt, class StateAt>
__device__ __forceinline__ bool elim_middle(int n_rows, int m, double add00, const BaseAt base_at,
                                            const StateAt state_at) {
    using D = Dim<O>;
    constexpr int B = D::B, ND = D::ND, NR = D::NR, NU = D::NU;
    const double *b = base_at(m);
    double d[ND], r[NR];
    row_load<BaseAt, D::F_D, ND>(b, d);
    row_load<BaseAt, D::F_R, NR>(b, r);
    d[0] += add00;
#pragma unroll
    for (int side = 0; side < 2; ++side) {  // 0: last row of the top half (m-1), 1: last row of the bottom half (m+1)
        if (side == 0 ? m > 0 : m < n_rows - 1) {
            const double *sn = state_at(side == 0 ? m - 1 : m + 1);
            double C[NU], W[NU], z[NR];
            load_coupling<O, BaseAt>(side == 0 ? base_at(m - 1) : b, side != 0, C);
            row_load<StateAt, D::SW, NU>(sn, W);
            row_load<StateAt, D::SX, NR>(sn, z);
#pragma unroll
            for (int p = 0; p < B; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) {
                    double acc = d[sym(p, q)];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], W[t * B + q], acc);
                    d[sym(p, q)] = acc;
                }
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double acc = r[p * 3 + x];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], z[t * 3 + x], acc);
                    r[p * 3 + x] = acc;
                }
            }
        }
    }
    double P[ND];
    const bool ok = sym_inverse<B>(d, P);
    double xm[NR];
#pragma unroll
    for (int p = 0; p < B; ++p)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double acc = P[sym(p, 0)] * r[x];
#pragma unroll
            for (int t = 1; t < B; ++t) acc = fma(P[sym(p, t)], r[t * 3 + x], acc);
            xm[p * 3 + x] = acc;
        }
    row_store<StateAt, D::SX, NR>(state_at(m), xm);
    return ok;
}

// One lane doing the whole elimination: top half, bottom half, split row.
template <int O, class BaseAt, class StateAt>
__device__ __forceinline__ bool thomas_forward(int n_rows, int m, double add00, const BaseAt base_at,
                                               const StateAt state_at) {
    if (n_rows <= 0) return true;
    bool ok = elim_half<O>(n_rows, m, false, add00, base_at, state_at);
    ok = elim_half<O>(n_rows, n_rows - 1 - m, true, add00, base_at, state_at) && ok;
    return elim_middle<O>(n_rows, m, add00, base_at, state_at) && ok;
}

// Where the back-substitution leaves the solution.  NoOut: nowhere (a speculative lane only needs its max deviation).
struct NoOut {
    static constexpr bool ENABLED = false;
    static constexpr bool VEC = false;
    static constexpr int FS = 1;
    using Mem = PlainMem;
    __device__ __forceinline__ double *operator()(int) const { return nullptr; }
    __device__ __forceinline__ void prefetch(int) const {}
};

// One back-substitution step: x_j = z_j - W_j x_n  (x_n = the solution of the neighbouring row nearer to the split row).
//   s: state row j (z, W) with field stride SFS;
//   xo: where x_j is stored ([r][axis], field stride XOut::FS) if XOut::ENABLED.
template <int O, class StateAt, class XOut>
__device__ __forceinline__ void thomas_back_step(const double *s, double *xo, const double (&xn)[3 * (O - 1)],
                                                 double (&x)[3 * (O - 1)]) {
    using D = Dim<O>;
    constexpr int B = D::B, NR = D::NR;
    double W[D::NU];
    row_load<StateAt, D::SX, NR>(s, x);
    row_load<StateAt, D::SW, D::NU>(s, W);
#pragma unroll
    for (int p = 0; p < B; ++p)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double acc = x[p * 3 + a];
#pragma unroll
            for (int q = 0; q < B; ++q) acc = fma(-W[p * B + q], xn[q * 3 + a], acc);
            x[p * 3 + a] = acc;
        }
    if constexpr (XOut::ENABLED) row_store<XOut, D::SX, NR>(xo, x);
}

// Squared deviation ratio of one segment (ms.cpp:594-617):  || p(t*) - L(t*) ||^2 / |P_{k+1} - P_k|^2.
// The segment is seen from one of its ends: the OUTER waypoint (position po, derivatives dout, [r-1][axis]) is the one
// farther from the split row, the INNER one (pi, din) the nearer.  outer_is_k1: the outer waypoint is the segment's
// second waypoint (bottom half), so its Hermite weights are the second half of h.  Terms are accumulated outer first.
// segx = { h[M], L[3], 1/len^2 } with compile-time stride XS.  The caller takes the square root of the maximum.
template <int O, int XS>
__device__ __forceinline__ double deviation_sq(const double *segx, bool outer_is_k1, const double (&po)[3],
                                               const double *dout, const double (&pi)[3], const double *din) {
    constexpr int M = 2 * O;
    const double *ho = segx + (outer_is_k1 ? O * XS : 0), *hi = segx + (outer_is_k1 ? 0 : O * XS);
    double d2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double p = ho[0] * po[a];
#pragma unroll
        for (int r = 1; r < O; ++r) p = fma(ho[r * XS], dout[(r - 1) * 3 + a], p);
        p = fma(hi[0], pi[a], p);
#pragma unroll
        for (int r = 1; r < O; ++r) p = fma(hi[r * XS], din[(r - 1) * 3 + a], p);
        const double dd = p - segx[(M + a) * XS];
        d2 = fma(dd, dd, d2);
    }
    return d2 * segx[(M + 3) * XS];
}

// Back-substitution of one half of a trajectory, outwards from the split row m, with the deviation probe of every
// segment it passes folded in (when EVAL).
//   mirror == false: rows m-1 .. 0, then segment 0 (its outer end is the fixed first waypoint, derivatives dfix = d0)
//   mirror == true : rows m+1 .. n-1, then the last segment (outer end: the fixed last waypoint, dfix = dN)
//   pos(w, out[3])  : position of waypoint w = 0..ns        (row j belongs to waypoint j+1)
//   state_at(j)     : state row j (z, W; SX of row m holds x_m);  xout(j): row that receives x_j, or NoOut
//   segx_at(k)      : deviation probe of segment k
// Returns the larger of m2_in and the squared deviation ratios seen (m2_in when !EVAL).
// EARLY: the caller only needs to know WHETHER the ratio exceeds the reweighting threshold 0.2 (ms.cpp:82), so the
// sweep stops at the first segment whose squared ratio is safely above 0.04 and returns that (partial) maximum --
// still > 0.2.  A sweep that runs to the end returns the exact maximum, so a trajectory that passes the test always
// has its exact max_dev.
constexpr double EARLY_DEV2 = 0.04 * (1.0 + 1e-9);
template <int O, bool EVAL, bool EARLY, class StateAt, class XOut, class SegxAt, class PosAt>
__device__ __forceinline__ double back_half(int n_rows, int m, bool mirror, const StateAt state_at, const XOut xout,
                                            const SegxAt segx_at, const PosAt pos, const double *dfix, double m2_in) {
    using D = Dim<O>;
    constexpr int NR = D::NR;
    double m2 = m2_in;
    double xn[NR], pn[3];  // solution and position of the row nearer to the split row
    {
        row_load<StateAt, D::SX, NR>(state_at(m), xn);
        if constexpr (XOut::ENABLED) {
            if (!mirror) row_store<XOut, D::SX, NR>(xout(m), xn);  // x_m itself (written once, by the top half)
        }
        pos(m + 1, pn);
    }
    const int cnt = mirror ? n_rows - 1 - m : m;
    const int step = mirror ? 1 : -1;
    int j = m + step;
    for (int i = 0; i < cnt; ++i, j += step) {
        if (i + 1 < cnt) state_at.prefetch(j + step);
        double x[NR], po[3];
        thomas_back_step<O, StateAt, XOut>(state_at(j), xout(j), xn, x);
        pos(j + 1, po);
        if (EVAL) {  // the segment between rows j and j -+ 1: segment j+1 (top half) or j (bottom half)
            if (EVAL) segx_at.prefetch(mirror ? j + 1 : j);
            m2 = fmax(m2, deviation_sq<O, SegxAt::FS>(segx_at(mirror ? j : j + 1), mirror, po, x, pn, xn));
        }
#pragma unroll
        for (int k = 0; k < NR; ++k) xn[k] = x[k];
        pn[0] = po[0]; pn[1] = po[1]; pn[2] = po[2];
        if (EVAL && EARLY && m2 > EARLY_DEV2) return m2;
    }
    if (EVAL) {  // the end segment: its outer waypoint has fixed derivatives
        double pe[3];
        pos(mirror ? n_rows + 1 : 0, pe);
        m2 = fmax(m2, deviation_sq<O, SegxAt::FS>(segx_at(mirror ? n_rows : 0), mirror, pe, dfix, pn, xn));
    }
    return m2;
}

// Squared deviation ratio of a single-segment trajectory (both ends fixed).
template <int O, class SegxAt, class PosAt>
__device__ __forceinline__ double single_segment_dev2(const SegxAt segx_at, const PosAt pos, const double *d0,
                                                      const double *dN) {
    double p0[3], p1[3];
    pos(0, p0);
    pos(1, p1);
    return deviation_sq<O, SegxAt::FS>(segx_at(0), false, p0, d0, p1, dN);
}

// One lane doing both halves: the bottom half first (for m = n-1 that is just the last segment, the likeliest to decide
// an EARLY test).  Returns max_k deviation ratio (0 when !EVAL).
template <int O, bool EVAL, bool EARLY, class StateAt, class XOut, class SegxAt, class PosAt>
__device__ __forceinline__ double thomas_backward(int n_rows, int m, const StateAt state_at, const XOut xout,
                                                  const SegxAt segx_at, const PosAt pos, const double *d0,
                                                  const double *dN) {
    if (n_rows <= 0) return EVAL ? sqrt(single_segment_dev2<O>(segx_at, pos, d0, dN)) : 0.0;
    double m2 = back_half<O, EVAL, EARLY>(n_rows, m, true, state_at, xout, segx_at, pos, dN, 0.0);
    if (EVAL && EARLY && m2 > EARLY_DEV2) return sqrt(m2);
    m2 = back_half<O, EVAL, EARLY>(n_rows, m, false, state_at, xout, segx_at, pos, d0, m2);
    return EVAL ? sqrt(m2) : 0.0;
}

// ---- one-axis forms of the twisted elimination ------------------------------------------------------------------
// The factorisation (D', P, W) of a chain is common to the three axes; the right-hand sides are not coupled.  A chain can
// therefore be walked by THREE lanes per half, each carrying the factorisation and ONE axis' right-hand side: per row 89
// FP64 instructions instead of 126 forward, 9 instead of 27 backward, on the critical path of a phase in which the other
// lanes of the CTA have nothing to do (the fused kernel's pass 1).  Every number is produced by the expression that produces
// it in elim_half / elim_middle / back_half, in the same order: the state rows are bit for bit those of the lane-pair form.
// `ax` = the lane's axis (0..2); all three lanes store the (identical) W.
template <class Rows>
__device__ __forceinline__ int field_off(int f) {
    if constexpr (RowPair<Rows>::LANES > 0) return (f >> 1) * (2 * RowPair<Rows>::LANES) + (f & 1);
    else return f * Rows::FS;
}

template <int O, class BaseAt, class StateAt>
__device__ __forceinline__ bool elim_half_ax(int n_rows, int cnt, bool mirror, double add00, int ax, const BaseAt base_at,
                                             const StateAt state_at) {
    using D = Dim<O>;
    constexpr int B = D::B, ND = D::ND, NU = D::NU;
    bool ok = true;
    double W[NU], z[B];
    const int step = mirror ? -1 : 1;
    int j = mirror ? n_rows - 1 : 0;
    auto row = [&](auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        const double *b = base_at(j);
        double *s = state_at(j);
        double d[ND], r[B];
        row_load<BaseAt, D::F_D, ND>(b, d);
#pragma unroll
        for (int p = 0; p < B; ++p) r[p] = b[field_off<BaseAt>(D::F_R + p * 3 + ax)];
        d[0] += add00;
        if constexpr (!FIRST) {
            double C[NU];
            load_coupling<O, BaseAt>(mirror ? b : base_at(j - 1), mirror, C);
#pragma unroll
            for (int p = 0; p < B; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) {
                    double acc = d[sym(p, q)];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], W[t * B + q], acc);
                    d[sym(p, q)] = acc;
                }
                double acc = r[p];
#pragma unroll
                for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], z[t], acc);
                r[p] = acc;
            }
        }
        double P[ND];
        ok = sym_inverse<B>(d, P) && ok;
#pragma unroll
        for (int p = 0; p < B; ++p) {
            double acc = P[sym(p, 0)] * r[0];
#pragma unroll
            for (int t = 1; t < B; ++t) acc = fma(P[sym(p, t)], r[t], acc);
            z[p] = acc;
        }
#pragma unroll
        for (int p = 0; p < B; ++p) s[field_off<StateAt>(D::SX + p * 3 + ax)] = z[p];
        {
            double N[NU];
            load_coupling<O, BaseAt>(mirror ? base_at(j - 1) : b, mirror, N);
#pragma unroll
            for (int p = 0; p < B; ++p)
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    double acc = P[sym(p, 0)] * N[q];
#pragma unroll
                    for (int t = 1; t < B; ++t) acc = fma(P[sym(p, t)], N[t * B + q], acc);
                    W[p * B + q] = acc;
                }
            row_store<StateAt, D::SW, NU>(s, W);
        }
        j += step;
    };
    if (cnt <= 0) return ok;
    row(std::true_type{});
    for (int i = 1; i < cnt; ++i) row(std::false_type{});
    return ok;
}

template <int O, class BaseAt, class StateAt>
__device__ __forceinline__ bool elim_middle_ax(int n_rows, int m, double add00, int ax, const BaseAt base_at,
                                               const StateAt state_at) {
    using D = Dim<O>;
    constexpr int B = D::B, ND = D::ND, NU = D::NU;
    const double *b = base_at(m);
    double d[ND], r[B];
    row_load<BaseAt, D::F_D, ND>(b, d);
#pragma unroll
    for (int p = 0; p < B; ++p) r[p] = b[field_off<BaseAt>(D::F_R + p * 3 + ax)];
    d[0] += add00;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? m > 0 : m < n_rows - 1) {
            const double *sn = state_at(side == 0 ? m - 1 : m + 1);
            double C[NU], W[NU], z[B];
            load_coupling<O, BaseAt>(side == 0 ? base_at(m - 1) : b, side != 0, C);
            row_load<StateAt, D::SW, NU>(sn, W);
#pragma unroll
            for (int p = 0; p < B; ++p) z[p] = sn[field_off<StateAt>(D::SX + p * 3 + ax)];
#pragma unroll
            for (int p = 0; p < B; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) {
                    double acc = d[sym(p, q)];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], W[t * B + q], acc);
                    d[sym(p, q)] = acc;
                }
                double acc = r[p];
#pragma unroll
                for (int t = 0; t < B; ++t) acc = fma(-C[t * B + p], z[t], acc);
                r[p] = acc;
            }
        }
    }
    double P[ND];
    const bool ok = sym_inverse<B>(d, P);
    double *sm = state_at(m);
#pragma unroll
    for (int p = 0; p < B; ++p) {
        double acc = P[sym(p, 0)] * r[0];
#pragma unroll
        for (int t = 1; t < B; ++t) acc = fma(P[sym(p, t)], r[t], acc);
        sm[field_off<StateAt>(D::SX + p * 3 + ax)] = acc;
    }
    return ok;
}

// back substitution of one half, one axis, solution left in the state rows (back_half<O, false, false> with xout = state_at)
template <int O, class StateAt>
__device__ __forceinline__ void back_half_ax(int n_rows, int m, bool mirror, int ax, const StateAt state_at) {
    using D = Dim<O>;
    constexpr int B = D::B, NU = D::NU;
    double xn[B];
    {
        const double *sm = state_at(m);
#pragma unroll
        for (int p = 0; p < B; ++p) xn[p] = sm[field_off<StateAt>(D::SX + p * 3 + ax)];
    }
    const int cnt = mirror ? n_rows - 1 - m : m;
    const int step = mirror ? 1 : -1;
    int j = m + step;
    for (int i = 0; i < cnt; ++i, j += step) {
        double *s = state_at(j);
        double x[B], W[NU];
#pragma unroll
        for (int p = 0; p < B; ++p) x[p] = s[field_off<StateAt>(D::SX + p * 3 + ax)];
        row_load<StateAt, D::SW, NU>(s, W);
#pragma unroll
        for (int p = 0; p < B; ++p) {
            double acc = x[p];
#pragma unroll
            for (int q = 0; q < B; ++q) acc = fma(-W[p * B + q], xn[q], acc);
            x[p] = acc;
        }
#pragma unroll
        for (int p = 0; p < B; ++p) {
            s[field_off<StateAt>(D::SX + p * 3 + ax)] = x[p];
            xn[p] = x[p];
        }
    }
}

// Polynomial coefficients of one segment and axis from its endpoint derivatives (c = M_k^-1 d, ms.cpp:584-591):
// out[i], i = 0..M-1, highest power first.  ip[e] = T^-e, pT[r] = T^r.
template <int O>
__device__ __forceinline__ void hermite_coeffs(const double (&yk)[O], const double (&yk1)[O], const double (&ip)[2 * O],
                                               const double (&pT)[O], double (&out)[2 * O]) {
    constexpr int M = 2 * O;
    double dh[M];
#pragma unroll
    for (int i = 0; i < O; ++i) {
        dh[i] = pT[i] * yk[i];
        dh[O + i] = pT[i] * yk1[i];
    }
#pragma unroll
    for (int k = 0; k < M; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) acc = fma(Tab<O>::H(k, i), dh[i], acc);
        out[M - 1 - k] = acc * ip[k];
    }
}

// Horner evaluation of one segment's xyz polynomials (coefficients highest power first; ms.cpp:104-117 evaluates
// the same polynomial as sum c*pow(t, e)).
template <int O>
__device__ __forceinline__ void eval_xyz(const double (&c)[3][2 * O], double t, double (&p)[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double v = c[a][0];
#pragma unroll
        for (int i = 1; i < 2 * O; ++i) v = fma(v, t, c[a][i]);
        p[a] = v;
    }
}

__device__ __forceinline__ double dist3(const double (&a)[3], const double (&b)[3]) {
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
}

// Acceptance test of the sampler (ms.cpp:143-145): || cur - prev || >= sample_distance.  The square root is only
// taken when the squared distance lies within 1e-14 (relative) of the squared threshold, where its rounding could
// matter; everywhere else comparing squares gives the same answer as the reference's sqrt-then-compare.
struct AcceptTest {
    double sd, lo, hi;
    __device__ __forceinline__ explicit AcceptTest(double sample_distance) : sd(sample_distance) {
        const double s2 = sample_distance > 0.0 ? sample_distance * sample_distance : 0.0;
        lo = s2 * (1.0 - 1e-14);
        hi = s2 * (1.0 + 1e-14);
    }
    // kept out of line so that the compiler cannot speculate the square root into the common path
    __device__ __noinline__ static bool tie_band(double d2, double sd_) { return sqrt(d2) >= sd_; }
    __device__ __forceinline__ bool operator()(const double (&a)[3], const double (&b)[3]) const {
        const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        if (d2 >= hi) return true;
        if (d2 < lo) return false;
        return tie_band(d2, sd);  // also reached for NaN (false) and for a non-positive threshold (true)
    }
};

}  // namespace msnap

#endif  // MSNAP_DEVICE_CUH
