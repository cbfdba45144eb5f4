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

// The library is a single translation unit (msnap_capi.cu), so the constant-bank copy of the tables is defined here.
__constant__ MsnapOrderTab c_tab[MSNAP_MAX_ORDER - MSNAP_MIN_ORDER + 1] = MSNAP_ORDER_TABLES;

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
    static constexpr int NSTATE = ND + NR;      // per-row factor + solution (order 4: 15 doubles)
    static constexpr int NSEGX = M + 4;         // per-segment deviation probe: h[M], L(t*)[3], 1/|P_{k+1}-P_k|
    static constexpr int NP = 2 * O - 1;        // inverse powers of T used by S: T^-1 .. T^-(2o-1)
};

__host__ __device__ constexpr int sym(int r, int q) { return r >= q ? r * (r + 1) / 2 + q : q * (q + 1) / 2 + r; }

// Tables of one order: compile-time indices -> constant-bank operands.
template <int O>
struct Tab {
    __device__ static __forceinline__ double S(int i, int j) { return c_tab[O - MSNAP_MIN_ORDER].S[i][j]; }
    __device__ static __forceinline__ double H(int k, int i) { return c_tab[O - MSNAP_MIN_ORDER].H[k][i]; }
    __device__ static __forceinline__ double HT(int s, int i) { return c_tab[O - MSNAP_MIN_ORDER].HT[s][i]; }
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
__device__ __forceinline__ void assemble_row(double Ta, double Tc, const double (&Pm)[3], const double (&P0)[3],
                                             const double (&Pp)[3], bool first, bool last, const Boundary<O> &bc,
                                             bool use_pw, double pw, int s_a, int s_c,
                                             const double *__restrict__ ht, double *row, int fs) {
    using D = Dim<O>;
    constexpr int B = D::B, NP = D::NP;
    double ipa[2 * O], ipc[2 * O], pTa[O], pTc[O];
    time_powers<O>(Ta, ipa, pTa);
    time_powers<O>(Tc, ipc, pTc);
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

// ---------------------------------------------------------------------------------------------------------
// Small dense kernels on (order-1) x (order-1) SPD blocks, everything in registers.
// ---------------------------------------------------------------------------------------------------------
// In-place Cholesky of a packed symmetric block: G lower, diagonal stored as RECIPROCAL (one division per pivot).
// Returns false if a pivot is not positive (or not finite).
template <int B>
__device__ __forceinline__ bool chol_packed(double (&g)[B * (B + 1) / 2]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < B; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = g[sym(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-g[sym(i, k)], g[sym(j, k)], s);
            if (i == j) {
                ok = ok && (s > 0.0);
                g[sym(i, i)] = rsqrt(s);  // 1 / G_ii
            } else {
                g[sym(i, j)] = s * g[sym(j, j)];
            }
        }
    }
    return ok;
}

// v <- G^-1 v  (forward substitution), v has stride `st`
template <int B>
__device__ __forceinline__ void fwd_solve(const double (&g)[B * (B + 1) / 2], double *v, int st) {
#pragma unroll
    for (int i = 0; i < B; ++i) {
        double s = v[i * st];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-g[sym(i, k)], v[k * st], s);
        v[i * st] = s * g[sym(i, i)];
    }
}

// v <- G^-T v  (backward substitution)
template <int B>
__device__ __forceinline__ void bwd_solve(const double (&g)[B * (B + 1) / 2], double *v, int st) {
#pragma unroll
    for (int i = B - 1; i >= 0; --i) {
        double s = v[i * st];
#pragma unroll
        for (int k = i + 1; k < B; ++k) s = fma(-g[sym(k, i)], v[k * st], s);
        v[i * st] = s * g[sym(i, i)];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Block-tridiagonal Cholesky solve of one trajectory for the three axes at once ("block Thomas").
//   base(j)  : read-only rows  (D, U, r),   state(j): per-row factor G (packed, reciprocal diagonal) and, after the
//   backward sweep, the solution x_j[r][axis] in the slot that held the forward-substituted right-hand side.
//   add00    : value added to D_j[0][0] of every row (2 * vel_zero_weight: one vw from each adjacent segment)
//   Storage is abstracted by two functors returning (pointer to field 0, field stride) for a row index.
// Returns false if any pivot failed.
// ---------------------------------------------------------------------------------------------------------
template <int O, class BaseAt, class StateAt>
__device__ __forceinline__ bool thomas_forward(int n_rows, double add00, BaseAt base_at, StateAt state_at) {
    using D = Dim<O>;
    constexpr int B = D::B, ND = D::ND, NR = D::NR;
    bool ok = true;
    double Y[B * B];  // Y = G_{j-1}^-1 U_{j-1}   (column q at Y[.*B + q])
    double w[NR];     // w = G_{j-1}^-1 r'_{j-1}, [r][axis]
    for (int j = 0; j < n_rows; ++j) {
        int bfs, sfs;
        const double *b = base_at(j, bfs);
        double *s = state_at(j, sfs);
        double g[ND], r[NR];
#pragma unroll
        for (int i = 0; i < ND; ++i) g[i] = b[(D::F_D + i) * bfs];
#pragma unroll
        for (int i = 0; i < NR; ++i) r[i] = b[(D::F_R + i) * bfs];
        g[0] += add00;
        if (j > 0) {
            // D'_j = D_j - Y'Y ;  r'_j = r_j - Y'w
#pragma unroll
            for (int p = 0; p < B; ++p) {
#pragma unroll
                for (int q = 0; q <= p; ++q) {
                    double acc = g[sym(p, q)];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-Y[t * B + p], Y[t * B + q], acc);
                    g[sym(p, q)] = acc;
                }
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    double acc = r[p * 3 + x];
#pragma unroll
                    for (int t = 0; t < B; ++t) acc = fma(-Y[t * B + p], w[t * 3 + x], acc);
                    r[p * 3 + x] = acc;
                }
            }
        }
        ok = chol_packed<B>(g) && ok;
#pragma unroll
        for (int x = 0; x < 3; ++x) fwd_solve<B>(g, r + x, 3);
#pragma unroll
        for (int i = 0; i < ND; ++i) s[i * sfs] = g[i];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            s[(ND + i) * sfs] = r[i];
            w[i] = r[i];
        }
        if (j + 1 < n_rows) {
#pragma unroll
            for (int i = 0; i < B * B; ++i) Y[i] = b[(D::F_U + i) * bfs];
#pragma unroll
            for (int q = 0; q < B; ++q) fwd_solve<B>(g, Y + q, B);
        }
    }
    return ok;
}

// One backward step: given x_{j+1} (xn, [r][axis]; ignored when `has_next` is false) produce x_j in place in
// state(j) and in x.
template <int O>
__device__ __forceinline__ void thomas_back_step(const double *b, int bfs, double *s, int sfs, bool has_next,
                                                 const double (&xn)[3 * (O - 1)], double (&x)[3 * (O - 1)]) {
    using D = Dim<O>;
    constexpr int B = D::B, ND = D::ND, NR = D::NR;
    double g[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) g[i] = s[i * sfs];
#pragma unroll
    for (int i = 0; i < NR; ++i) x[i] = s[(ND + i) * sfs];
    if (has_next) {
        double t[NR];
#pragma unroll
        for (int p = 0; p < B; ++p) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < B; ++q) acc = fma(b[(D::F_U + p * B + q) * bfs], xn[q * 3 + a], acc);
                t[p * 3 + a] = acc;
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) fwd_solve<B>(g, t + a, 3);
#pragma unroll
        for (int i = 0; i < NR; ++i) x[i] -= t[i];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) bwd_solve<B>(g, x + a, 3);
#pragma unroll
    for (int i = 0; i < NR; ++i) s[(ND + i) * sfs] = x[i];
}

// Deviation probe of one segment (ms.cpp:594-617): || p(t*) - L(t*) || / |P_{k+1} - P_k| from the endpoint
// derivatives.  segx = { h[M], L[3], rlen } with stride xs.  yk/yk1: [axis][o] endpoint derivative vectors.
template <int O>
__device__ __forceinline__ double deviation_ratio(const double *segx, int xs, const double (&yk)[3][O],
                                                  const double (&yk1)[3][O]) {
    constexpr int M = 2 * O;
    double d2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double p = 0.0;
#pragma unroll
        for (int i = 0; i < O; ++i) p = fma(segx[i * xs], yk[a][i], p);
#pragma unroll
        for (int i = 0; i < O; ++i) p = fma(segx[(O + i) * xs], yk1[a][i], p);
        const double dd = p - segx[(M + a) * xs];
        d2 = fma(dd, dd, d2);
    }
    return sqrt(d2) * segx[(M + 3) * xs];
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

}  // namespace msnap

#endif  // MSNAP_DEVICE_CUH
