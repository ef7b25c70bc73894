#!/usr/bin/env python
"""NumPy prototype of the device eigensolver (design evidence, not product, not oracle).

Same structure as csrc/eig_dc.cu: Householder tridiagonalisation (one fused update+symv pass per
step), Cuppen divide & conquer on the tridiagonal (rank-sort, LAPACK-style deflation, secular
roots by bisection on the bit pattern of the offset from the nearer pole, Gu-Eisenstat z-hat,
GEMM), reflector back-transformation.  Used to settle formulas, deflation tolerances and the
bisection bracket before the CUDA port; checked against LAPACK here on the CPU.

  python scripts/proto_eig_dc.py [n]
"""
import sys

import numpy as np

EPS = np.finfo(np.float64).eps / 2   # unit roundoff 1.1e-16 (LAPACK dlamch('E'))


# ---------------------------------------------------------------------------------------------
def tridiagonalize(A):
    """A (symmetric, full storage) -> d, e, V (reflector vectors in columns, v[j+1]=1), tau."""
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n); e = np.zeros(max(n - 1, 0)); tau = np.zeros(max(n - 1, 0))
    V = np.zeros((n, n))
    for j in range(n - 2):
        x = A[j + 1:, j].copy()
        alpha = x[0]
        xnorm = np.linalg.norm(x[1:])
        if xnorm == 0.0:
            tau[j] = 0.0; e[j] = alpha; d[j] = A[j, j]
            V[j + 1, j] = 1.0
            continue
        beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
        t = (beta - alpha) / beta
        v = x / (alpha - beta); v[0] = 1.0
        tau[j] = t; e[j] = beta; d[j] = A[j, j]
        V[j + 1:, j] = v
        A22 = A[j + 1:, j + 1:]
        p = t * (A22 @ v)
        w = p - (0.5 * t * (p @ v)) * v
        A22 -= np.outer(v, w) + np.outer(w, v)
    if n >= 2:
        d[n - 2] = A[n - 2, n - 2]; e[n - 2] = A[n - 1, n - 2]
    d[n - 1] = A[n - 1, n - 1]
    return d, e, V, tau


def backtransform(V, tau, Z):
    """U = H_0 H_1 ... H_{n-3} Z, applied reflector by reflector (columns of Z independent)."""
    U = Z.copy()
    n = V.shape[0]
    for j in range(n - 3, -1, -1):
        if tau[j] == 0.0:
            continue
        v = V[j + 1:, j]
        s = tau[j] * (v @ U[j + 1:, :])
        U[j + 1:, :] -= np.outer(v, s)
    return U


# ---------------------------------------------------------------------------------------------
def tql_leaf(d, e):
    """Implicit QL with Wilkinson shift (EISPACK tql2 / NR tqli flavour) for a small tridiagonal."""
    n = d.size
    d = d.copy(); e = np.append(e.copy(), 0.0)
    Z = np.eye(n)
    for l in range(n):
        it = 0
        while True:
            m = l
            while m < n - 1:
                dd = abs(d[m]) + abs(d[m + 1])
                if abs(e[m]) <= EPS * dd:
                    break
                m += 1
            if m == l:
                break
            it += 1
            if it > 60:
                raise RuntimeError("tql: no convergence")
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = np.hypot(g, 1.0)
            g = d[m] - d[l] + e[l] / (g + np.copysign(r, g))
            s = c = 1.0; p = 0.0
            i = m - 1
            broke = False
            while i >= l:
                f = s * e[i]; b = c * e[i]
                r = np.hypot(f, g); e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= p; e[m] = 0.0; broke = True
                    break
                s = f / r; c = g / r
                g = d[i + 1] - p
                r = (d[i] - g) * s + 2.0 * c * b
                p = s * r; d[i + 1] = g + p; g = c * r - b
                zi1 = Z[:, i + 1].copy()
                Z[:, i + 1] = s * Z[:, i] + c * zi1
                Z[:, i] = c * Z[:, i] - s * zi1
                i -= 1
            if broke:
                continue
            d[l] -= p; e[l] = g; e[m] = 0.0
    return d, Z


# ---------------------------------------------------------------------------------------------
def f2i(x):
    return np.float64(x).view(np.int64)


def i2f(i):
    return np.int64(i).view(np.float64)


def secular_root(j, dk, z2, rho):
    """Root j of 1 + rho*sum z2_i/(dk_i - lam) in (dk_j, dk_{j+1}) (last: (dk_k, dk_k + rho*sum z2)).
    Returns (origin index, mu) with lam = dk[origin] + mu.  Bisection on the bit pattern of |mu|."""
    k = dk.size

    def feval(org, mu):
        return 1.0 + rho * np.sum(z2 / ((dk - dk[org]) - mu))

    if j < k - 1:
        gap = dk[j + 1] - dk[j]
        half = 0.5 * gap
        fmid = feval(j, half)
        if fmid > 0.0:          # root in the left half: origin d_j, mu in (0, half]
            org, sign, hi = j, 1.0, half
        else:                   # root in the right half: origin d_{j+1}, mu in [-half, 0)
            org, sign, hi = j + 1, -1.0, half
    else:
        org, sign = k - 1, 1.0
        hi = rho * np.sum(z2)
        # f(hi) >= 0 in exact arithmetic; widen a little for rounding
        hi = hi * (1.0 + 8 * EPS) + 1e-300
    # invariant: g(lo) < 0 <= g(hi) with g(a) = sign * f(org, sign*a) increasing in a ... careful:
    # sign=+1: f increasing in mu, f(0+) = -inf, f(hi) > 0.
    # sign=-1: mu = -a; f(-a) decreasing in a; f(0-) = +inf, f(-half) <= 0.  Use h(a) = -f(-a): increasing,
    #          h(0+) = -inf, h(half) >= 0.
    lo_i = 0
    hi_i = int(f2i(hi))
    while hi_i - lo_i > 1:
        mid_i = lo_i + (hi_i - lo_i) // 2      # (lo+hi)//2 would overflow int64 on the device for |mu| >= 2
        a = float(i2f(mid_i))
        val = sign * feval(org, sign * a)
        if val < 0.0:
            lo_i = mid_i
        else:
            hi_i = mid_i
    a = float(i2f(hi_i))
    return org, sign * a


def merge(d1, Q1, d2, Q2, beta, stats=None):
    """Eigen-decomposition of [T1' 0; 0 T2'] + |beta| u u^T given those of T1', T2'."""
    n1, n2 = d1.size, d2.size
    n = n1 + n2
    s = 1.0 if beta >= 0 else -1.0
    rho = 2.0 * abs(beta)
    z = np.concatenate([Q1[-1, :], s * Q2[0, :]]) / np.sqrt(2.0)
    d = np.concatenate([d1, d2])
    Q = np.zeros((n, n))
    Q[:n1, :n1] = Q1
    Q[n1:, n1:] = Q2
    # rank sort ascending (ties by index)
    order = np.argsort(d, kind="stable")
    d = d[order]; z = z[order]; Q = Q[:, order]
    dmax = np.abs(d).max(); zmax = np.abs(z).max()
    tol = 8.0 * EPS * max(dmax, zmax)
    if rho * zmax <= tol:
        return d, Q
    # deflation scan (dlaed2 logic on the sorted sequence)
    defl = np.zeros(n, dtype=bool)
    prev = -1
    for i in range(n):
        if rho * abs(z[i]) <= tol:
            defl[i] = True
            continue
        if prev < 0:
            prev = i
            continue
        # try to deflate prev against i
        sv, cv = -z[prev], z[i]
        tau = np.hypot(cv, sv)
        t = d[i] - d[prev]
        cv /= tau; sv /= tau
        if abs(t * cv * sv) <= tol:
            z[i] = tau; z[prev] = 0.0
            qp = Q[:, prev].copy(); qi = Q[:, i].copy()
            Q[:, prev] = cv * qp + sv * qi
            Q[:, i] = -sv * qp + cv * qi
            tt = d[prev] * cv * cv + d[i] * sv * sv
            d[i] = d[prev] * sv * sv + d[i] * cv * cv
            d[prev] = tt
            defl[prev] = True
            # keep sortedness of the deflated value irrelevant (parent re-sorts)
            prev = i
        else:
            prev = i
    nd = np.nonzero(~defl)[0]
    k = nd.size
    if stats is not None:
        stats.append((n, k))
    dk = d[nd]; zk = z[nd]
    # after rotations dk stays ascending? d[i] >= old d[i]?  (c^2 d_i + s^2 d_prev is between them)  -> re-check
    assert np.all(np.diff(dk) > 0), "non-deflated poles must be strictly increasing"
    z2 = zk * zk
    org = np.zeros(k, dtype=np.int64); mu = np.zeros(k)
    for j in range(k):
        org[j], mu[j] = secular_root(j, dk, z2, rho)
    # Gu-Eisenstat z-hat:  zh_i^2 = prod_j (lam_j - d_i) / (rho * prod_{j != i} (d_j - d_i))
    # lam_j - d_i = (dk[org_j] - dk[i]) + mu_j
    L = (dk[org][None, :] - dk[:, None]) + mu[None, :]          # L[i, j] = lam_j - d_i
    Dm = dk[None, :] - dk[:, None]                              # D[i, j] = d_j - d_i
    np.fill_diagonal(Dm, 1.0)
    ratio = L / Dm                                              # diag = lam_i - d_i
    zh = np.sqrt(np.prod(ratio, axis=1) / rho) * np.sign(zk)
    S = zh[:, None] / (-L)                                      # S[i, j] = zh_i / (d_i - lam_j)
    S /= np.linalg.norm(S, axis=0)[None, :]
    lam = dk[org] + mu
    Qn = np.empty((n, n)); dn = np.empty(n)
    Qn[:, :k] = Q[:, nd] @ S
    dn[:k] = lam
    Qn[:, k:] = Q[:, defl]
    dn[k:] = d[defl]
    return dn, Qn


def dc_tridiag(d, e, leaf=32, stats=None):
    n = d.size
    if n <= leaf:
        return tql_leaf(d, e)
    m = n // 2
    beta = e[m - 1]
    d1 = d[:m].copy(); d2 = d[m:].copy()
    d1[-1] -= abs(beta); d2[0] -= abs(beta)
    l1, Q1 = dc_tridiag(d1, e[:m - 1], leaf, stats)
    l2, Q2 = dc_tridiag(d2, e[m:], leaf, stats)
    return merge(l1, Q1, l2, Q2, beta, stats)


def sym_eig(A, leaf=32, stats=None):
    d, e, V, tau = tridiagonalize(A)
    lam, Z = dc_tridiag(d, e, leaf, stats)
    U = backtransform(V, tau, Z)
    o = np.argsort(-lam, kind="stable")
    return lam[o], U[:, o]


def check(A, name, leaf=32):
    n = A.shape[0]
    stats = []
    lam, U = sym_eig(A, leaf, stats)
    ref = np.linalg.eigvalsh(A)[::-1]
    nrm = max(np.abs(ref).max(), 1e-300)
    err = np.abs(lam - ref).max() / nrm
    orth = np.abs(U.T @ U - np.eye(n)).max()
    res = np.abs(A @ U - U * lam[None, :]).max() / nrm
    print(f"{name:28s} n={n:5d} eig rel err {err:.2e} orth {orth:.2e} resid {res:.2e}  "
          f"count(1e-10) {(lam >= 1e-10).sum()} vs {(ref >= 1e-10).sum()}  merges(n,k) top: {stats[-1] if stats else None}")
    assert err < 1e-13 * max(1, n / 100) and orth < 1e-12 and res < 1e-12, name


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for n in (1, 2, 3, 5, 33, 64, 65, 100, nmax):
        A = rng.standard_normal((n, n)); A = (A + A.T) / 2
        check(A, "random symmetric")
    check(np.eye(70), "identity")
    check(np.zeros((70, 70)), "zeros")
    check(np.ones((70, 70)), "ones (rank 1)")
    B = rng.standard_normal((90, 7)); check(B @ B.T, "rank 7 PSD")
    Qr, _ = np.linalg.qr(rng.standard_normal((120, 120)))
    check(Qr @ np.diag(np.repeat([3.0, 1.0, -2.0, 0.0], 30)) @ Qr.T, "4 clusters x30")
    check(Qr @ np.diag(np.logspace(2, -17, 120)) @ Qr.T, "graded 1e2..1e-17")
    T = np.diag(2.0 * np.ones(150)) - np.diag(np.ones(149), 1) - np.diag(np.ones(149), -1)
    check(T, "1-2-1 tridiagonal")
    W = np.diag(np.abs(np.arange(-40, 41)).astype(float)) + np.diag(np.ones(80), 1) + np.diag(np.ones(80), -1)
    check(W, "Wilkinson W81+")
    sys.path.insert(0, __file__.rsplit("/", 2)[0])
    try:
        from bench import synth_luminance
        from oracle import nle_oracle as O
        g = int(np.sqrt(nmax))
        lum = synth_luminance(1024, 1024).astype(np.float64)
        sel, _ = O.sample_pixels(1024, 1024, g, g)
        Ka = O.affinity_block(lum.ravel(), 1024, sel, sel, 500.0, 30.0)
        check(Ka, f"Ka bench image {g}x{g}")
    except ImportError as ex:
        print("skipping Ka:", ex)
