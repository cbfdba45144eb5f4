"""The arithmetic k_alt_part runs (csrc/msnap_alt_part.cuh), restated in numpy and checked against a
dense solve: an SPD pentadiagonal system cut into P chunks with 2-row separators; every chunk's interior eliminated by the
downward LDL' recurrence with two coupling columns V = L^-1 H[chunk, separator above] and continued into its separator
below; the separators' 2 x 2 block-tridiagonal system eliminated from both ends; back substitution through the chunks.
No GPU: this pins the derivation (which Schur complement goes where, the transposes of the couplings, the split of the
rows over the lanes), the GPU tests pin the kernel."""
import numpy as np
import pytest
from scipy.linalg import cho_factor, cho_solve


def geometry(n, gw):
    """Rows of lane p = [start, start + cnt) interior + 2 separator rows (all but the last lane) -- alt_part_set's rule."""
    P = min(n // 4, gw) if n >= 8 else 1
    m = n - 2 * (P - 1)
    bs, rem = divmod(m, P)
    return [(p * (bs + 2) + min(p, rem), bs + (1 if p < rem else 0), p < P - 1) for p in range(P)]


def inv2(A):
    """Inverse of a symmetric 2 x 2 block by cofactors (the kernel's formulas: no pivoting, which on a block with one
    1e8 penalty row would subtract two 1e11-sized numbers)."""
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[0, 1]
    return np.array([[A[1, 1], -A[0, 1]], [-A[0, 1], A[0, 0]]]) / det


def partitioned_solve(d, e, h, b, gw):
    """H = diag(d) + off-diagonals e (k, k+1) and h (k, k+2); returns z with H z = b."""
    n = len(d)
    chunks = geometry(n, gw)
    P = len(chunks)
    fac, A, g, C = [], [], [], []
    for p, (start, cnt, hr) in enumerate(chunks):
        hl = p > 0
        a1 = a2 = cc = Dm1 = Dm2 = ym1 = ym2 = 0.0
        v1m1 = v1m2 = v2m1 = v2m2 = 0.0
        H1, H2, H2n = (h[start - 2], e[start - 1], h[start - 1]) if hl else (0.0, 0.0, 0.0)
        q11 = q12 = q22 = r1 = r2 = 0.0
        rows = []
        for i in range(cnt):
            k = start + i
            D = d[k] - a1 * a1 * Dm1 - a2 * a2 * Dm2
            y = b[k] - a1 * ym1 - a2 * ym2
            v1 = H1 - a1 * v1m1 - a2 * v1m2
            v2 = H2 - a1 * v2m1 - a2 * v2m2
            H1, H2, H2n = 0.0, H2n, 0.0
            assert D > 0
            ek = e[k] if k + 1 < n else 0.0
            hk = h[k] if k + 2 < n else 0.0
            n1, n2 = (ek - cc * Dm1 * a1) / D, hk / D
            rows.append((a1, a2, y / D, v1 / D, v2 / D))
            q11 += v1 * v1 / D; q12 += v1 * v2 / D; q22 += v2 * v2 / D; r1 += v1 * y / D; r2 += v2 * y / D
            a2, a1, cc = cc, n1, n2
            Dm2, Dm1, ym2, ym1 = Dm1, D, ym1, y
            v1m2, v1m1, v2m2, v2m1 = v1m1, v1, v2m1, v2
        fac.append((rows, (a1, a2, cc), (q11, q12, q22, r1, r2)))
        if hr:   # the separator below: this chunk's eliminations applied, not eliminated itself
            k = start + cnt
            A.append(np.array([[d[k] - a1 * a1 * Dm1 - a2 * a2 * Dm2, e[k] - cc * Dm1 * a1],
                               [e[k] - cc * Dm1 * a1, d[k + 1] - cc * cc * Dm1]]))
            g.append(np.array([b[k] - a1 * ym1 - a2 * ym2, b[k + 1] - cc * ym1]))
            C.append(np.array([[-(a1 * v1m1 + a2 * v1m2), -(a1 * v2m1 + a2 * v2m2)], [-cc * v1m1, -cc * v2m1]]))
    nb = P - 1
    for t in range(nb):   # ... and the contribution of the chunk below it
        q11, q12, q22, r1, r2 = fac[t + 1][2]
        A[t] = A[t] - np.array([[q11, q12], [q12, q22]])
        g[t] = g[t] - np.array([r1, r2])
    s = [None] * nb
    if nb:
        mb = nb // 2
        for t in range(1, mb):                    # from the top: coupling C_t to block t-1
            M = C[t] @ inv2(A[t - 1]); A[t] = A[t] - M @ C[t].T; g[t] = g[t] - M @ g[t - 1]
        for t in range(nb - 2, mb, -1):           # from the bottom: coupling C_{t+1}' to block t+1
            M = C[t + 1].T @ inv2(A[t + 1]); A[t] = A[t] - M @ C[t + 1]; g[t] = g[t] - M @ g[t + 1]
        if mb >= 1:
            M = C[mb] @ inv2(A[mb - 1]); A[mb] = A[mb] - M @ C[mb].T; g[mb] = g[mb] - M @ g[mb - 1]
        if mb + 1 <= nb - 1:
            M = C[mb + 1].T @ inv2(A[mb + 1]); A[mb] = A[mb] - M @ C[mb + 1]; g[mb] = g[mb] - M @ g[mb + 1]
        s[mb] = inv2(A[mb]) @ g[mb]
        for t in range(mb - 1, -1, -1):
            s[t] = inv2(A[t]) @ (g[t] - C[t + 1].T @ s[t + 1])
        for t in range(mb + 1, nb):
            s[t] = inv2(A[t]) @ (g[t] - C[t] @ s[t - 1])
    z = np.zeros(n)
    for p, (start, cnt, hr) in enumerate(chunks):
        rows, (a1, a2, cc), _ = fac[p]
        sL = s[p - 1] if p > 0 else np.zeros(2)
        z1 = z2 = b1 = b2 = b2n = 0.0
        if hr:
            z[start + cnt], z[start + cnt + 1] = s[p]
            z1, z2, b1, b2, b2n = s[p][0], s[p][1], a1, cc, a2
        for i in range(cnt - 1, -1, -1):
            l1, l2, yd, vd1, vd2 = rows[i]
            zk = yd - vd1 * sL[0] - vd2 * sL[1] - b1 * z1 - b2 * z2
            z[start + i] = zk
            z2, z1, b2, b2n, b1 = z1, zk, b2n, l2, l1
    return z


@pytest.mark.parametrize("gw", [16, 32, 8])
def test_partitioned_elimination_equals_the_dense_solve(gw):
    rng = np.random.default_rng(5 + gw)
    for n in list(range(1, 41)) + [63, 64, 65, 150, 205, 259, 270, 271, 400, 542]:
        # the altitude system's shape: smoothing stencil (second differences), climb weights, a few 1e8 / 1e10 penalty rows
        L = np.zeros((max(n - 2, 0), n))
        for i in range(n - 2):
            L[i, i:i + 3] = [1.0, -2.0, 1.0]
        w = rng.uniform(0.01, 0.3, n)
        Hm = 10.0 * L.T @ L + 1e-8 * np.eye(n)
        for k in range(n - 1):
            Hm[k, k] += w[k]; Hm[k + 1, k + 1] += w[k]; Hm[k, k + 1] -= w[k]; Hm[k + 1, k] -= w[k]
        follow = np.where(rng.uniform(size=n) < 0.9, 1.0, 0.0)          # pass 1's follow term on the rows the map covers
        pen = np.where(rng.uniform(size=n) < 0.3, 1e8, 0.0)              # pass 2's active-set rows and pinned ends
        pen[0] += 1e10; pen[-1] += 1e10
        Hm += np.diag(follow + pen)
        b = (follow + pen) * rng.uniform(1200, 1400, n)
        d = np.diag(Hm).copy()
        e = np.append(np.diag(Hm, 1), 0.0) if n > 1 else np.zeros(1)
        h = np.append(np.diag(Hm, 2), [0.0, 0.0])[:n] if n > 2 else np.zeros(n)
        want = cho_solve(cho_factor(Hm), b)   # (a symmetric factorisation: LU with row pivoting is itself ~1e-6 m off here)
        got = partitioned_solve(d, e, h, b, gw)
        # backward stable like any elimination of an SPD matrix: a residual at rounding level (row-wise scaled: the penalty
        # rows are 1e10 times heavier than the others) ...
        scale = np.abs(Hm) @ np.abs(got) + np.abs(b)
        assert (np.abs(Hm @ got - b) / scale).max() <= 1e-13, (n, gw)
        # ... and a dense Cholesky solve's heights far below the north star's 1e-6 m
        assert np.abs(got - want).max() <= 1e-9, (n, gw)
