"""Opt-in GPU test of the shared-memory-resident tridiagonalisation (NLE_B200_TRD=resident, eig_dc.cu).

The kernel is an experiment that measured within 3 % of the default (profiles/r1l_trd_phases.md), so it stays OFF
by default and this file is skipped unless NLE_B200_TEST_TRD_RESIDENT=1.  It performs the same arithmetic in the
same order as tridiag_kernel, so the eigen-decomposition must come out BIT-identical with the switch on and off
(observed on the device for n = 3, 4, 5, 33, 149, 300, 612, 1041, 1600, 1800 and identity / zero / all-ones /
diagonal matrices).  The kernel spins on flagged cells: give the process a limit, e.g.

    NLE_B200_TEST_TRD_RESIDENT=1 timeout 300 python -m pytest tests/test_gpu_trd_resident.py -x -q
"""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NLE_B200_TEST_TRD_RESIDENT") != "1",
                                 reason="experimental kernel: set NLE_B200_TEST_TRD_RESIDENT=1")]


def both(nb, A, **kw):
    os.environ.pop("NLE_B200_TRD", None)
    U0, D0 = nb.eigenDecomposition(A, **kw)
    os.environ["NLE_B200_TRD"] = "resident"
    os.environ["NLE_B200_EIG_STRICT"] = "1"      # no silent Jacobi fallback if the sanity check trips
    try:
        U1, D1 = nb.eigenDecomposition(A, **kw)
    finally:
        os.environ.pop("NLE_B200_TRD", None)
        os.environ.pop("NLE_B200_EIG_STRICT", None)
    return (U0, D0), (U1, D1)


@pytest.mark.parametrize("n", [3, 4, 5, 33, 130, 147, 148, 149, 300, 700, 1041, 1600, 1800])
def test_resident_equals_grid_sync_bitwise(nb, n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, max(3, n // 2)))
    A = B @ B.T / n + 1e-3 * np.eye(n)
    (U0, D0), (U1, D1) = both(nb, A, eps=-1e300)
    assert np.array_equal(D0, D1)
    assert np.array_equal(U0, U1)
    w = np.linalg.eigvalsh(A)[::-1]
    assert np.abs(D1 - w).max() <= 1e-11 * max(1.0, np.abs(w).max()) * max(1, n / 16)


def test_resident_zero_tail_and_identity(nb):
    for A in (np.eye(70), np.zeros((70, 70)), np.ones((70, 70)), np.diag(np.arange(1.0, 41.0))):
        (U0, D0), (U1, D1) = both(nb, A, eps=-1e300)
        assert np.array_equal(D0, D1) and np.array_equal(U0, U1)


def test_resident_train_equals_default(nb):
    from nle_testlib import synth_lum
    lum = synth_lum(160, 200, seed=3)
    args = (12, 14, 60.0, 25.0, 8, 12)
    os.environ.pop("NLE_B200_TRD", None)
    f0 = nb.NLEFilter().trainFilter(lum, *args)
    os.environ["NLE_B200_TRD"] = "resident"
    try:
        f1 = nb.NLEFilter().trainFilter(lum, *args)
    finally:
        os.environ.pop("NLE_B200_TRD", None)
    assert np.array_equal(f0.eigvals, f1.eigvals)
    w = [2.0, 3.0, 4.0, 1.0]
    assert np.array_equal(f0.enhanceLuminance(lum, w), f1.enhanceLuminance(lum, w))


@pytest.mark.parametrize("env", [{"NLE_B200_TRD_DYN": "4"}, {"NLE_B200_TRD_DYN": "16"}, {"NLE_B200_TRD_GRID": "37"},
                                 {"NLE_B200_TRD": "resident", "NLE_B200_TRD_GRID": "100"},
                                 {"NLE_B200_TRD": "cluster", "NLE_B200_TRD_CLUSTER": "2"},
                                 {"NLE_B200_TRD": "cluster", "NLE_B200_TRD_CLUSTER": "4"},
                                 {"NLE_B200_TRD": "cluster", "NLE_B200_TRD_CLUSTER": "8"}])
@pytest.mark.parametrize("n", [5, 64, 149, 700, 1041, 1600])
def test_grid_knobs_are_bit_identical(nb, env, n):
    """NLE_B200_TRD_DYN / NLE_B200_TRD_GRID only change which CTA owns a column, never a column's arithmetic."""
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, max(3, n // 2)))
    A = B @ B.T / n + 1e-3 * np.eye(n)
    for k in ("NLE_B200_TRD", "NLE_B200_TRD_DYN", "NLE_B200_TRD_GRID", "NLE_B200_TRD_CLUSTER"):
        os.environ.pop(k, None)
    U0, D0 = nb.eigenDecomposition(A, eps=-1e300)
    os.environ.update(env)
    os.environ["NLE_B200_EIG_STRICT"] = "1"
    try:
        U1, D1 = nb.eigenDecomposition(A, eps=-1e300)
    finally:
        for k in list(env) + ["NLE_B200_EIG_STRICT"]:
            os.environ.pop(k, None)
    assert np.array_equal(D0, D1) and np.array_equal(U0, U1)
