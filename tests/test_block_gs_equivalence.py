"""The Bullet-order kernel (csrc/snake_pgs.cu) sweeps its 112 rows in blocks of <= 32 (16 motor rows | 32 normal rows | 2 x 16
friction pairs): w = J dv for all rows of a block at once, the rows in order with the in-block Delassus entries A(i, j) = J_j . B_i
carrying every impulse change to the later rows, dv updated once per block.  This numpy restatement of both forms (same row order,
clamps and cone projection as oracle/snake_oracle.c `tick`) checks that the block form produces the SAME iterates as the row-by-row
projected Gauss-Seidel -- the rearrangement is exact, not a Jacobi-style relaxation."""
import numpy as np

NJ, NC, ND = 16, 32, 22
MU = 2.0


def _system(rng, inactive=()):
    L = rng.normal(size=(ND, ND))
    Minv = L @ L.T / ND + np.eye(ND)
    J = np.zeros((NJ + 3 * NC, ND))
    for j in range(NJ):
        J[j, 6 + j] = 1.0
    J[NJ:] = rng.normal(size=(3 * NC, ND))
    for c in inactive:  # a separated contact has all-zero rows in the kernel
        J[NJ + c] = 0.0; J[NJ + NC + 2 * c] = 0.0; J[NJ + NC + 2 * c + 1] = 0.0
    B = J @ Minv
    D = np.einsum("rk,rk->r", J, B)
    invD = np.where(D > 0, 1.0 / np.where(D > 0, D, 1.0), 0.0)
    rhs = rng.normal(size=len(J)) * invD
    return J, B, invD, rhs


def _cone(sa, sb, lim):
    n2 = sa * sa + sb * sb
    if n2 > lim * lim:
        sc = lim / np.sqrt(n2)
        return sa * sc, sb * sc
    return sa, sb


def rowwise(J, B, invD, rhs, iters, maximp):
    lam = np.zeros(len(J)); dv = np.zeros(ND)
    for _ in range(iters):
        for j in range(NJ):
            d = rhs[j] - (J[j] @ dv) * invD[j]
            s = np.clip(lam[j] + d, -maximp, maximp)
            d = s - lam[j]; lam[j] = s; dv += B[j] * d
        for c in range(NC):
            r = NJ + c
            d = rhs[r] - (J[r] @ dv) * invD[r]
            s = max(lam[r] + d, 0.0)
            d = s - lam[r]; lam[r] = s; dv += B[r] * d
        for c in range(NC):
            ra = NJ + NC + 2 * c; rb = ra + 1
            lim = MU * lam[NJ + c]
            sa = lam[ra] + rhs[ra] - (J[ra] @ dv) * invD[ra]
            sb = lam[rb] + rhs[rb] - (J[rb] @ dv) * invD[rb]
            sa, sb = _cone(sa, sb, lim)
            da, db = sa - lam[ra], sb - lam[rb]
            lam[ra], lam[rb] = sa, sb
            dv += B[ra] * da + B[rb] * db
    return lam, dv


def blocked(J, B, invD, rhs, iters, maximp):
    lam = np.zeros(len(J)); dv = np.zeros(ND)
    A = J @ B.T  # A[j, i] = J_j . B_i
    blocks = [(0, NJ, "motor"), (NJ, NC, "normal"), (NJ + NC, 32, "friction"), (NJ + NC + 32, 32, "friction")]
    for _ in range(iters):
        for base, m, kind in blocks:
            rows = np.arange(base, base + m)
            w = J[rows] @ dv
            dl = np.zeros(m)
            step = 2 if kind == "friction" else 1
            for i in range(0, m, step):
                r = base + i
                if kind == "motor":
                    s = np.clip(lam[r] + rhs[r] - w[i] * invD[r], -maximp, maximp)
                    dl[i] = s - lam[r]; lam[r] = s
                elif kind == "normal":
                    s = max(lam[r] + rhs[r] - w[i] * invD[r], 0.0)
                    dl[i] = s - lam[r]; lam[r] = s
                else:
                    c = (r - NJ - NC) // 2
                    lim = MU * lam[NJ + c]
                    sa = lam[r] + rhs[r] - w[i] * invD[r]
                    sb = lam[r + 1] + rhs[r + 1] - w[i + 1] * invD[r + 1]
                    sa, sb = _cone(sa, sb, lim)
                    dl[i], dl[i + 1] = sa - lam[r], sb - lam[r + 1]
                    lam[r], lam[r + 1] = sa, sb
                for q in range(step):  # the later rows of the block see the change through A
                    w[i + step:] += A[rows[i + step:], r + q] * dl[i + q]
            dv += B[rows].T @ dl
    return lam, dv


def test_block_form_reproduces_rowwise_iterates():
    rng = np.random.default_rng(0)
    for trial in range(6):
        inactive = () if trial % 2 == 0 else tuple(rng.choice(NC, 5, replace=False))
        maximp = np.inf if trial < 3 else 0.3
        J, B, invD, rhs = _system(rng, inactive)
        for iters in (1, 3, 12):
            l0, v0 = rowwise(J, B, invD, rhs, iters, maximp)
            l1, v1 = blocked(J, B, invD, rhs, iters, maximp)
            assert np.allclose(l0, l1, rtol=1e-9, atol=1e-11), (trial, iters, np.abs(l0 - l1).max())
            assert np.allclose(v0, v1, rtol=1e-9, atol=1e-11)
        assert (l0[NJ:NJ + NC] > 0).any() and (l0[NJ:NJ + NC] == 0).any()  # the clamps are exercised
