"""GPU: every N-scaled stage has two hand-written implementations -- the cell / level-table kernels that ship
(cell_kernels.cu, sinkhorn_cells.cu) and the pixel-axis / per-row kernels they replaced (filter_kernels.cu), selected by
an environment variable read once per process.  Both are re-associations of the same FP64 sums, so their
stage outputs must agree to rounding and the enhanced image must be identical.  Each variant runs in its own process."""
import os
import subprocess
import sys

import numpy as np
import pytest

from nle_testlib import ROOT

pytestmark = pytest.mark.gpu

CASES = [(96, 128, 8, 10, 40.0, 25.0, 6, 8), (120, 333, 20, 10, 300.0, 30.0, 5, 10), (160, 150, 50, 50, 200.0, 10.0, 3, 20),
         (64, 48, 3, 1, 20.0, 15.0, 3, 2), (256, 320, 24, 17, 60.0, 12.0, 4, 60)]

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[2] + '/tests')
import nonlocal_image_edit_b200 as nb
from nle_testlib import synth_lum
nb.load().nle_b200_set_keep_stages(1)
res = {}
for i, (h, w, a, b, hx, hy, T, k) in enumerate(%r):
    L = synth_lum(h, w, seed=11 + i)
    f = nb.NLEFilter().trainFilter(L, a, b, hx, hy, T, k)
    for st, nm in ((2, 'rvec'), (3, 'c'), (7, 'G')):
        res[f'{nm}{i}'] = f.stage(st)
    res[f'S{i}'] = f.eigvals
    res[f'V{i}'] = f.eigvecs
    res[f'out{i}'] = f.enhanceLuminance(L, [2.0, 3.0, 4.0, 1.0])
np.savez(sys.argv[1], **res)
""" % (CASES,)


def _run(tmp_path, tag, env_extra):
    env = {k: v for k, v in os.environ.items() if not k.startswith("NLE_B200_")}
    env.update(env_extra)
    out = str(tmp_path / f"{tag}.npz")
    subprocess.run([sys.executable, "-c", CHILD, out, ROOT], env=env, check=True, timeout=600)
    return np.load(out)


@pytest.fixture(scope="module")
def default_run(tmp_path_factory):
    return _run(tmp_path_factory.mktemp("paths"), "default", {})


@pytest.mark.parametrize("var,val", [("NLE_B200_GRAM", "pixel"), ("NLE_B200_SINKHORN", "rows"), ("NLE_B200_EXT", "pixel"),
                                     ("NLE_B200_SK_UNFUSED", "1"), ("NLE_B200_SK_STAGED", "1")])
def test_alternate_kernels_agree(tmp_path, default_run, var, val):
    alt = _run(tmp_path, "alt", {var: val})
    for i in range(len(CASES)):
        for nm, tol in (("rvec", 1e-9), ("c", 1e-8), ("G", 1e-9), ("S", 1e-9)):
            a, b = alt[f"{nm}{i}"], default_run[f"{nm}{i}"]
            assert a.shape == b.shape
            assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300), (var, i, nm)
        Va, Vb = alt[f"V{i}"], default_run[f"V{i}"]
        assert Va.shape == Vb.shape
        for j in range(Vb.shape[1]):
            s = np.sign(np.dot(Va[:, j], Vb[:, j])) or 1.0
            assert np.abs(s * Va[:, j] - Vb[:, j]).max() <= 1e-6 * max(np.abs(Vb[:, j]).max(), 1e-300), (var, i, j)
        d = np.abs(alt[f"out{i}"].astype(int) - default_run[f"out{i}"].astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3, (var, i)
