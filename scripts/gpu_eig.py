"""Developer diagnostic: eigensolver timing / accuracy on Ka matrices of the bench workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb
from oracle import nle_oracle as O
from bench import synth_luminance
lum = synth_luminance(1024, 1024).astype(np.float64)
sel, _ = O.sample_pixels(1024, 1024, 40, 40)
Kfull = O.affinity_block(lum.ravel(), 1024, sel, sel, 500.0, 30.0)
sizes = [int(x) for x in (sys.argv[1:] or [200, 592, 600, 608, 1031, 1600])]
for n in sizes:
    idx = np.linspace(0, 1599, n).astype(int)
    A = Kfull[np.ix_(idx, idx)]
    w = np.linalg.eigvalsh(A)[::-1]
    nb.eigenDecomposition(A[:16, :16])
    t0 = time.time(); U, D = nb.eigenDecomposition(A, eps=-1e300); dt = time.time() - t0
    print(f'n={n} inner={os.environ.get("NLE_B200_EIG_INNER","2")} {dt*1e3:.1f} ms eig abs err {np.abs(D - w).max():.2e} count {(D >= 1e-10).sum()} vs {(w >= 1e-10).sum()} orth {np.abs(U.T @ U - np.eye(n)).max():.1e} resid {np.abs(A @ U - U * D).max():.1e}', flush=True)
    if os.environ.get('NLE_SAVE'):
        res = np.abs(A @ U - U * D).max(axis=0)
        np.savez(os.path.join(ROOT, 'gpurun_out', f'eig_n{n}.npz'), D=D, w=w, res=res)
