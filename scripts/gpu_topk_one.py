"""One call of the top-k block solver (for `ncu --metrics gpu__time_duration.sum`): python scripts/gpu_topk_one.py n k decay"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nonlocal_image_edit_b200 as nb
n, k, decay = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
rng = np.random.default_rng(1)
lam = np.concatenate([[1.0, 0.86, 0.72], 0.6 * decay ** np.arange(n - 3)]) + 1e-9
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
A = (Q * lam) @ Q.T
A = (A + A.T) / 2
U, D, prod = nb.topkEigenDecomposition(A, k, assume_psd=True, return_products=True)
print("products", prod, "pairs", D.size)
