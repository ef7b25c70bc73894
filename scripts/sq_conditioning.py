#!/usr/bin/env python
"""How well-conditioned are the small eigenvalues Sq?  Evidence behind the tolerance of tests/test_gpu_parity.py::sq_close.

north_star asks for "eigenvalues within 1e-5 relative".  This script evaluates the REFERENCE algebra twice in FP64 -- the dense
restatement (oracle.train_dense, line by line filter.cpp:480-502) and the factor-form restatement (oracle.train_streaming, the
same mathematics re-associated) -- on README images and reports the relative difference of every eigenvalue Sq_i against its
size relative to Sq_0.  Two FP64 evaluation orders of the same formula bound what ANY faithful implementation can promise:
where they differ by more than 1e-5 relative, the reference itself does not define the eigenvalue to 1e-5.

  python scripts/sq_conditioning.py [image ...]        # default: brickwall forest bird paper; writes profiles/sq_conditioning.md
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nle_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    names = sys.argv[1:] or ["brickwall", "forest", "bird", "paper"]
    man = {m["name"]: m for m in json.load(open(os.path.join(GOLDEN, "manifest.json")))}
    lines = ["# Conditioning of the eigenvalues Sq (scripts/sq_conditioning.py)", "",
             "Dense FP64 restatement vs factor-form FP64 restatement of the same reference algebra (both in oracle/nle_oracle.py).",
             "`rel` = |Sq_dense − Sq_stream| / Sq_dense.  Buckets by Sq_i / Sq_0.", "",
             "| image | k' | r / r2 | Sq range | max rel, Sq_i ≥ 1e-4·Sq_0 | max rel, Sq_i < 1e-4·Sq_0 | smallest Sq_i and its rel |", "|---|---|---|---|---|---|---|"]
    for name in names:
        m = man[name]
        img = cv2.imread(os.path.join(GOLDEN, f"{name}_input.png"))
        lum = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)[:, :, 0].astype(np.float64)
        args = (m["n_row_samples"], m["n_col_samples"], m["hx"], m["hy"], m["n_sinkhorn_iter"], m["n_eigen_vectors"])
        fd = O.train_dense(lum, *args)
        fs = O.train_streaming(lum, *args, block_fn=O.affinity_block_c)
        assert (fd.stages["r"], fd.stages["r2"], fd.eigvals.size) == (fs.stages["r"], fs.stages["r2"], fs.eigvals.size)
        S, T = fd.eigvals, fs.eigvals
        rel = np.abs(S - T) / S
        big = S >= 1e-4 * S[0]
        a = f"{rel[big].max():.1e}" if big.any() else "–"
        b = f"{rel[~big].max():.1e}" if (~big).any() else "– (none)"
        lines.append(f"| {name} | {S.size} | {fd.stages['r']} / {fd.stages['r2']} | {S[0]:.4f} … {S[-1]:.2e} | {a} | {b} | {S[-1]:.2e}: {rel[-1]:.1e} |")
        print(lines[-1], file=sys.stderr, flush=True)
    lines += ["", "Reading: two FP64 evaluation orders of the reference's algebra agree to better than 1e-6 relative on EVERY eigenvalue, including",
              "the ones far below 1e-4 of the largest (brickwall uses the whole positive block of Wa, Sq down to ~7e-6: 5e-7).  Round 1's",
              "`sq_close` relaxed the bound below that floor on the strength of an uncommitted experiment; this table does not support the",
              "relaxation, so the tests now assert north_star's 1e-5 relative on every eigenvalue (tests/test_gpu_parity.py::sq_close)."]
    open(os.path.join(ROOT, "profiles", "sq_conditioning.md"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
