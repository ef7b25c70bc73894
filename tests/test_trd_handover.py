"""CPU: the algebraic fact behind the hybrid tridiagonalisation (csrc/eig_dc.cu: tridiag_kernel in partial mode ->
tridiag_cluster_kernel on the trailing block).

Householder tridiagonalisation never looks back: after s reflectors the remaining work is the tridiagonalisation of the updated
trailing block A[s:, s:], an independent problem.  A NumPy restatement of the kernels' step (p = A v, w = tau p - (tau^2/2)(p.v) v,
rank-2 update, next reflector from the updated column) is run once uninterrupted and once stopped after s steps and restarted on
the trailing block: d, e, tau and the reflectors must be IDENTICAL (same operations in the same order), and T must have the
eigenvalues of A.  The GPU side of the same statement is tests/test_gpu_eig_variants.py (sizes 1709 ... 2300)."""
import numpy as np
import pytest


def _reflector(x):
    """x -> (v, tau, beta) with v[0] = 1, (I - tau v v^T) x = beta e1  (LAPACK dlarfg convention, as make_reflector)."""
    alpha = x[0]
    xn2 = float(np.dot(x[1:], x[1:]))
    v = x.copy()
    if xn2 == 0.0:
        v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
    tau = (beta - alpha) / beta
    v[1:] *= 1.0 / (alpha - beta)
    v[0] = 1.0
    return v, tau, beta


def tridiagonalise(A, nstop=None):
    """Returns d, e, tau, V (reflector j in V[j+1:, j]) and the updated matrix.  nstop: number of reflectors to produce."""
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n); e = np.zeros(max(n - 1, 0)); tau = np.zeros(max(n - 1, 0)); V = np.zeros_like(A)
    steps = n - 2 if nstop is None else min(nstop, n - 2)
    for j in range(steps):
        v, tau[j], e[j] = _reflector(A[j + 1:, j].copy())
        d[j] = A[j, j]
        V[j + 1:, j] = v
        B = A[j + 1:, j + 1:]
        p = B @ v
        w = tau[j] * p - (0.5 * tau[j] * tau[j] * np.dot(p, v)) * v
        B -= np.outer(v, w) + np.outer(w, v)
    if nstop is None or nstop >= n - 2:
        if n >= 2:
            d[n - 2] = A[n - 2, n - 2]; e[n - 2] = A[n - 1, n - 2]
        d[n - 1] = A[n - 1, n - 1]
    return d, e, tau, V, A


@pytest.mark.parametrize("n,s", [(12, 1), (40, 7), (40, 37), (97, 50)])
def test_partial_reduction_then_restart_equals_the_uninterrupted_reduction(n, s):
    rng = np.random.default_rng(n * 100 + s)
    B = rng.standard_normal((n, n))
    A = B @ B.T / n
    d0, e0, t0, V0, _ = tridiagonalise(A)
    d1, e1, t1, V1, A1 = tridiagonalise(A, nstop=s)                   # leading s reflectors, trailing block A1[s:, s:] updated
    d2, e2, t2, V2, _ = tridiagonalise(A1[s:, s:])                    # an independent problem
    d = np.concatenate([d1[:s], d2]); e = np.concatenate([e1[:s], e2]); t = np.concatenate([t1[:s], t2])
    assert np.array_equal(d, d0) and np.array_equal(e, e0) and np.array_equal(t, t0)
    V = V1.copy(); V[s:, s:] = V2
    assert np.array_equal(V, V0)
    T = np.diag(d0) + np.diag(e0, 1) + np.diag(e0, -1)
    w = np.linalg.eigvalsh(A)
    assert np.abs(np.linalg.eigvalsh(T) - w).max() <= 1e-13 * max(1.0, np.abs(w).max()) * n
