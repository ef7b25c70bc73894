"""-m gpu: topkEigenDecomposition (filter.cpp:169-200) through the C ABI -- the Chebyshev-filtered block solver that the training
path uses for eig(Q) (csrc/eig_topk.cu) against LAPACK, and the Spectra selection rules of the reference on the full-solver route."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def psd(n, lam, seed=0):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * lam) @ Q.T
    return (A + A.T) / 2


@pytest.mark.parametrize("n,k,decay", [(700, 50, 0.96), (1357, 50, 0.985), (1300, 100, 0.985), (420, 20, 0.9)])
def test_block_solver_matches_lapack(nb, n, k, decay):
    lam = np.concatenate([[1.0, 0.86, 0.72], 0.6 * decay ** np.arange(n - 3)]) + 1e-9
    A = psd(n, lam, seed=n)
    U, D, products = nb.topkEigenDecomposition(A, k, assume_psd=True, return_products=True)
    assert products > 0                                        # the block solver ran and converged
    w = np.linalg.eigvalsh(A)[::-1]
    assert D.size == k and U.shape == (n, k)
    assert np.abs(D - w[:k]).max() <= 1e-13 * w[0]
    assert np.abs(A @ U - U * D).max() <= 1e-12 * w[0]
    assert np.abs(U.T @ U - np.eye(k)).max() <= 1e-12


def test_block_solver_handles_multiple_eigenvalues(nb):
    n, k = 500, 24
    lam = np.concatenate([[2.0] * 3, [1.5] * 5, 0.8 * 0.93 ** np.arange(n - 8)])
    A = psd(n, lam, seed=3)
    U, D, products = nb.topkEigenDecomposition(A, k, assume_psd=True, return_products=True)
    assert products > 0
    w = np.sort(lam)[::-1]
    assert np.abs(D - w[:k]).max() <= 1e-13 * w[0]
    assert np.abs(A @ U - U * D).max() <= 1e-12 * w[0]


def test_block_solver_cuts_at_eps(nb):
    n, k = 520, 30
    lam = np.concatenate([np.linspace(1.0, 0.1, 12), np.full(n - 12, 1e-13)])        # rank 12: the 13th eigenvalue is below eps
    A = psd(n, lam, seed=5)
    U, D, products = nb.topkEigenDecomposition(A, k, assume_psd=True, return_products=True)
    assert products <= 0                                        # gave up (k-th eigenvalue < eps) -> full solver, same contract
    assert D.size == 12 and np.allclose(D, lam[:12], rtol=1e-12)


@pytest.mark.parametrize("n,k", [(60, 5), (612, 50)])
def test_blocks_below_the_break_even_size_take_the_full_solver(nb, n, k):
    lam = np.linspace(2.0, 0.1, n)
    A = psd(n, lam, seed=7)
    U, D, products = nb.topkEigenDecomposition(A, k, assume_psd=True, return_products=True)
    assert products == 0
    assert np.allclose(D, lam[:k], rtol=1e-12)
    assert np.abs(A @ U - U * D).max() <= 1e-12 * lam[0]


def test_spectra_selection_rules(nb):
    """nev = min(nLargest, n-1) (:172); LARGEST_MAGN selection reported in descending algebraic order; prefix cut at eps."""
    lam = np.array([3.0, 1.0, 0.5, -0.2, -4.0])
    A = psd(5, lam, seed=9)
    U, D = nb.topkEigenDecomposition(A, 10, eps=-1e300)
    assert D.size == 4                                          # n - 1
    assert np.allclose(D, [3.0, 1.0, 0.5, -4.0], atol=1e-12)    # |.|-largest four, algebraic order; -0.2 dropped
    U, D = nb.topkEigenDecomposition(A, 2, eps=-1e300)
    assert np.allclose(D, [3.0, -4.0], atol=1e-12)
    U, D = nb.topkEigenDecomposition(A, 2)                      # default eps = 1e-10: prefix rule stops at -4
    assert np.allclose(D, [3.0], atol=1e-12)
    assert np.abs(A @ U - U * D).max() <= 1e-12
    with pytest.raises(nb.NleError):
        nb.topkEigenDecomposition(np.ones((1, 1)), 1)
