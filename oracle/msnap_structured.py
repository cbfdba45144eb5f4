"""oracle/msnap_structured.py -- TEST INFRASTRUCTURE, not product code.

"Oracle B" of SURVEY.md section 8(c): the same optimisation problem as
``TrajectoryGeneratorTool::SolveQPClosedForm`` (minimum_snap.cpp:227-649), restated in the *structured* form
(per-segment 2o x 2o Hermite-space Hessians, block-tridiagonal reduced system on the free interior derivatives)
and evaluated either in IEEE double (``ctx=None``) or in multi-precision arithmetic (``ctx=mpmath.mp`` with
``mp.dps >= 50``).  In multi-precision it is the *truth* against which both the reference's dense-LU result and the
GPU result are measured (the reference's own fp64 error is not negligible at order >= 4); in double it documents
the algorithm the CUDA kernels implement.  Pure-Python loops: small cases only.

Formulation ("equivalent structured form"):
  segment k, duration T, endpoint-derivative vector d = [y_k ; y_{k+1}],  y_j = (p, p', ..., p^(o-1)) at waypoint j
  cost_k = d' S_k d + g_k' d,     S_k = M_k^-T (Q_k + pw*phi phi' + vw*V_k) M_k^-1,    g_k = M_k^-T f_k
    M_k^-T Q_k M_k^-1   = T^(1-2o) * D S_hat D,  D = diag(T^(i mod o))      (ms.cpp:247-266, 313-330)
    M_k^-T phi(t*)      = h = D H(t*/T)  (Hermite basis values)              (ms.cpp:441-446)
    M_k^-T V_k M_k^-1   = e_1 e_1' + e_{o+1} e_{o+1}'                        (ms.cpp:474-509)
    g_k = -2 pw L(t*) h   (NOT halved in the stationarity condition, ms.cpp:579)
  unknowns x_j = y_j[1:], j = 1..ns-1:   L_j x_{j-1} + D_j x_j + U_j x_{j+1} = rhs_j
"""
from __future__ import annotations

from fractions import Fraction
from functools import lru_cache
from math import factorial


# ------------------------------------------------------------------------------------------------------------
# exact constant tables per order (rationals)
# ------------------------------------------------------------------------------------------------------------
def _mat_inv_frac(A):
    n = len(A)
    M = [list(map(Fraction, row)) + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(A)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        piv = M[c][c]
        M[c] = [v / piv for v in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                f = M[r][c]
                M[r] = [a - f * b for a, b in zip(M[r], M[c])]
    return [row[n:] for row in M]


@lru_cache(maxsize=None)
def tables(order: int):
    """Exact tables on the unit interval, ascending powers tau^k:
       H[k][i]  : coefficient of tau^k in the Hermite basis polynomial of endpoint derivative i  (= M_hat^-1)
       S[i][j]  : Hessian of  int_0^1 (d^o p / dtau^o)^2  in Hermite space                       (= H' Q_hat H)
       HT[s][i] : basis polynomial i evaluated at tau = s/16, s = 0..16                           (ms.cpp:408-416)"""
    o, m = order, 2 * order
    Mh = [[Fraction(0)] * m for _ in range(m)]
    for j in range(o):
        for k in range(j, m):
            r = Fraction(factorial(k), factorial(k - j))
            Mh[j][k] = r if k == j else Fraction(0)      # tau = 0
            Mh[o + j][k] = r                                # tau = 1
    H = _mat_inv_frac(Mh)
    Q = [[Fraction(0)] * m for _ in range(m)]
    for k in range(o, m):
        for l in range(o, m):
            Q[k][l] = Fraction(factorial(k), factorial(k - o)) * Fraction(factorial(l), factorial(l - o)) / (k + l - 2 * o + 1)
    S = [[sum(H[k][i] * Q[k][l] * H[l][j] for k in range(m) for l in range(m)) for j in range(m)] for i in range(m)]
    HT = [[sum(H[k][i] * Fraction(s, 16) ** k for k in range(m)) for i in range(m)] for s in range(17)]
    return H, S, HT


# ------------------------------------------------------------------------------------------------------------
# number contexts
# ------------------------------------------------------------------------------------------------------------
class _F64:
    @staticmethod
    def num(x):
        return float(x) if not isinstance(x, Fraction) else x.numerator / x.denominator

    @staticmethod
    def sqrt(x):
        return x ** 0.5


class _MP:
    def __init__(self, mp):
        self.mp = mp

    def num(self, x):
        if isinstance(x, Fraction):
            return self.mp.mpf(x.numerator) / self.mp.mpf(x.denominator)
        return self.mp.mpf(float(x)) if not isinstance(x, self.mp.mpf) else x

    def sqrt(self, x):
        return self.mp.sqrt(x)


def _ctx(ctx):
    return _F64 if ctx is None else _MP(ctx)


def _chol_solve(A, B, c):
    """Solve A X = B for SPD A (n x n) and B (n x r) by Cholesky; returns X."""
    n = len(A)
    if n == 0:
        return []
    Lm = [[c.num(0)] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1):
            s = A[i][j] - sum(Lm[i][k] * Lm[j][k] for k in range(j))
            Lm[i][j] = c.sqrt(s) if i == j else s / Lm[j][j]
    r = len(B[0])
    X = [[c.num(0)] * r for _ in range(n)]
    for col in range(r):
        y = [c.num(0)] * n
        for i in range(n):
            y[i] = (B[i][col] - sum(Lm[i][k] * y[k] for k in range(i))) / Lm[i][i]
        for i in reversed(range(n)):
            X[i][col] = (y[i] - sum(Lm[k][i] * X[k][col] for k in range(i + 1, n))) / Lm[i][i]
    return X


# ------------------------------------------------------------------------------------------------------------
# the solve
# ------------------------------------------------------------------------------------------------------------
def solve_structured(order, Path, Vel, Acc, Time, path_weight=0.0, vel_zero_weight=0.0, ctx=None, best_s=None):
    """Structured equivalent of SolveQPClosedForm.  Returns dict(coeff[ns][3][m] highest power first,
    max_dev, best_s[ns], dist2[ns][17], ratio[ns], y[ns+1][3][o]).  ``best_s`` forces the arg-max decisions
    (used to compare arithmetic only, independent of tie-breaking)."""
    c = _ctx(ctx)
    o, m, b = order, 2 * order, order - 1
    ns = len(Time)
    Hf, Sf, HTf = tables(order)
    H = [[c.num(v) for v in row] for row in Hf]
    Sh = [[c.num(v) for v in row] for row in Sf]
    HT = [[c.num(v) for v in row] for row in HTf]
    P = [[c.num(Path[j][a]) for a in range(3)] for j in range(ns + 1)]
    T = [c.num(t) for t in Time]
    pw, vw = c.num(path_weight), c.num(vel_zero_weight)
    zero = c.num(0)

    # fixed derivative values y_j[a][r]; free ones are filled by the solve
    y = [[[zero] * o for _ in range(3)] for _ in range(ns + 1)]
    for j in range(ns + 1):
        for a in range(3):
            y[j][a][0] = P[j][a]
    for a in range(3):
        if o >= 2:
            y[0][a][1] = c.num(Vel[0][a]); y[ns][a][1] = c.num(Vel[1][a])
        if o >= 3:
            y[0][a][2] = c.num(Acc[0][a]); y[ns][a][2] = c.num(Acc[1][a])

    def seg_matrix(k, use_pw, use_vw, s_star):
        """S_k (m x m) and g_k[a] (m)."""
        Tk = T[k]
        Dg = [Tk ** (i % o) for i in range(m)]
        scale = Tk ** (1 - 2 * o)
        S = [[scale * Dg[i] * Sh[i][j] * Dg[j] for j in range(m)] for i in range(m)]
        g = [[zero] * m for _ in range(3)]
        if use_pw:
            h = [Dg[i] * HT[s_star[k]][i] for i in range(m)]
            tau = c.num(Fraction(s_star[k], 16))
            for i in range(m):
                for j in range(m):
                    S[i][j] = S[i][j] + pw * h[i] * h[j]
            for a in range(3):
                La = P[k][a] + tau * (P[k + 1][a] - P[k][a])
                for i in range(m):
                    g[a][i] = c.num(-2) * pw * La * h[i]
        if use_vw:
            S[1][1] = S[1][1] + vw
            S[o + 1][o + 1] = S[o + 1][o + 1] + vw
        return S, g

    def solve_pass(use_pw, use_vw, s_star):
        n = ns - 1
        if n > 0 and b > 0:
            segs = [seg_matrix(k, use_pw, use_vw, s_star) for k in range(ns)]
            Dm = [[[zero] * b for _ in range(b)] for _ in range(n)]
            Um = [[[zero] * b for _ in range(b)] for _ in range(n)]
            rhs = [[[zero] * 3 for _ in range(b)] for _ in range(n)]       # [row][r][axis]
            for jj in range(n):
                j = jj + 1
                Sa, ga = segs[j - 1]
                Sb, gb = segs[j]
                for r in range(b):
                    for q in range(b):
                        Dm[jj][r][q] = Sa[o + 1 + r][o + 1 + q] + Sb[1 + r][1 + q]
                        Um[jj][r][q] = Sb[1 + r][o + 1 + q]
                    for a in range(3):
                        acc = Sa[o + 1 + r][0] * y[j - 1][a][0] + Sa[o + 1 + r][o] * y[j][a][0]
                        acc = acc + Sb[1 + r][0] * y[j][a][0] + Sb[1 + r][o] * y[j + 1][a][0]
                        if j - 1 == 0:
                            for q in range(1, o):
                                acc = acc + Sa[o + 1 + r][q] * y[0][a][q]
                        if j + 1 == ns:
                            for q in range(1, o):
                                acc = acc + Sb[1 + r][o + q] * y[ns][a][q]
                        acc = acc + ga[a][o + 1 + r] + gb[a][1 + r]
                        rhs[jj][r][a] = -acc
            # block Thomas: forward elimination
            Dp = [None] * n
            Rp = [None] * n
            W = [None] * n
            for jj in range(n):
                Dj = [row[:] for row in Dm[jj]]
                Rj = [row[:] for row in rhs[jj]]
                if jj > 0:
                    # L_j = U_{j-1}'  ;  D'_j = D_j - L_j W_{j-1},  r'_j = r_j - L_j z_{j-1}
                    for r in range(b):
                        for q in range(b):
                            Dj[r][q] = Dj[r][q] - sum(Um[jj - 1][t][r] * W[jj - 1][0][t][q] for t in range(b))
                        for a in range(3):
                            Rj[r][a] = Rj[r][a] - sum(Um[jj - 1][t][r] * W[jj - 1][1][t][a] for t in range(b))
                WU = _chol_solve(Dj, Um[jj], c)     # D'^-1 U_j
                Z = _chol_solve(Dj, Rj, c)          # D'^-1 r'_j
                W[jj] = (WU, Z)
            X = [None] * n
            for jj in reversed(range(n)):
                WU, Z = W[jj]
                X[jj] = [[Z[r][a] - (sum(WU[r][q] * X[jj + 1][q][a] for q in range(b)) if jj + 1 < n else zero)
                          for a in range(3)] for r in range(b)]
            for jj in range(n):
                for r in range(b):
                    for a in range(3):
                        y[jj + 1][a][1 + r] = X[jj][r][a]

    def hermite_eval(k, s, a):
        Tk = T[k]
        return sum((Tk ** (i % o)) * HT[s][i] * (y[k][a][i] if i < o else y[k + 1][a][i - o]) for i in range(m))

    dist2 = [[zero] * 17 for _ in range(ns)]
    s_star = [0] * ns
    if path_weight > 0:
        solve_pass(False, False, None)                              # ms.cpp:349-405
        for k in range(ns):
            best = c.num(-1)
            for s in range(17):
                tau = c.num(Fraction(s, 16))
                d2 = zero
                for a in range(3):
                    La = P[k][a] + tau * (P[k + 1][a] - P[k][a])
                    dd = hermite_eval(k, s, a) - La
                    d2 = d2 + dd * dd
                dist2[k][s] = d2
                if d2 > best:
                    best, s_star[k] = d2, s
        if best_s is not None:
            s_star = list(best_s)
    solve_pass(path_weight > 0, vel_zero_weight > 0, s_star)       # ms.cpp:468-592

    ratios = [zero] * ns
    max_dev = zero
    for k in range(ns):                                             # ms.cpp:594-624
        tau = c.num(Fraction(s_star[k], 16))
        d2 = zero
        sl = zero
        for a in range(3):
            La = P[k][a] + tau * (P[k + 1][a] - P[k][a])
            dd = hermite_eval(k, s_star[k], a) - La
            d2 = d2 + dd * dd
            sl = sl + (P[k + 1][a] - P[k][a]) ** 2
        seg_len = c.sqrt(sl)
        ratios[k] = c.sqrt(d2) / seg_len if seg_len > 1e-6 else zero
        if ratios[k] > max_dev:
            max_dev = ratios[k]

    coeff = [[[zero] * m for _ in range(3)] for _ in range(ns)]
    for k in range(ns):
        Tk = T[k]
        for a in range(3):
            dh = [(Tk ** (i % o)) * (y[k][a][i] if i < o else y[k + 1][a][i - o]) for i in range(m)]
            for p in range(m):
                coeff[k][a][m - 1 - p] = sum(H[p][i] * dh[i] for i in range(m)) / (Tk ** p)
    return dict(coeff=coeff, max_dev=max_dev, best_s=s_star, dist2=dist2, ratio=ratios, y=y)


def reweighted_structured(order, Path, Vel, Acc, Time, path_weight, vel_zero_weight, ctx=None, forced=None):
    """The reweighting loop of ms.cpp:76-90 around ``solve_structured``.
    ``forced`` = (iters, best_s) replays the reference's discrete decisions (arithmetic-only comparison)."""
    it, vw = 0, vel_zero_weight
    while True:
        out = solve_structured(order, Path, Vel, Acc, Time, path_weight, vw, ctx, None if forced is None else forced[1])
        go_on = (out["max_dev"] > 0.2 and it < 10) if forced is None else (it < forced[0])
        if go_on:
            vw = 0.01 if vw < 1e-6 else vw * 2.0
            it += 1
        else:
            break
    out["iters"] = it
    out["vw_final"] = vw
    return out
