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
              "relaxation, so the tests assert north_star's 1e-5 relative on every eigenvalue (tests/test_gpu_parity.py::sq_close); the one",
              "input in the suite where the reference's algebra does not pin Sq to 1e-5 is in the second table."]
    lines += sweep_table()
    open(os.path.join(ROOT, "profiles", "sq_conditioning.md"), "w").write("\n".join(lines) + "\n")


def sweep_table():
    """Second table: the hx / hy sweep of tests/test_gpu_parity.py::test_hx_hy_sweep_matches_oracle (72 x 88 synthetic image,
    9 x 11 samples, T = 6, k = 12).  The dense restatement is evaluated with LAPACK's MRRR (dsyevr, scipy's and the oracle's default), QR
    iteration (dsyev, the family of Eigen's SelfAdjointEigenSolver) and divide & conquer (dsyevd) eigensolvers, and in factor form."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scipy.linalg
    from nle_testlib import oracle_sq_spread, synth_lum
    orig = scipy.linalg.eigh
    out = ["", "## hx / hy sweep: the same algebra under three LAPACK eigensolvers and in factor form", "",
           "`rel` = max_i |Sq_i(variant) − Sq_i(dense, dsyevr)| / Sq_i.  72×88 synthetic image, 9×11 samples, T = 6, k = 12.", "",
           "| hx | hy | r / r2 | smallest Sq | dsyev (QR iteration) | dsyevd (divide & conquer) | factor form | all + 8 × (Ka + E), ‖E‖₂ = ε‖Ka‖₂ |", "|---|---|---|---|---|---|---|---|"]
    worst = 0.0
    for hx in (5.0, 100.0, 5000.0):
        for hy in (3.0, 10.0, 30.0, 100.0):
            L = synth_lum(72, 88, seed=21).astype(np.float64)
            a = (9, 11, hx, hy, 6, 12)
            base = O.train_dense(L, *a)
            rel = []
            try:
                for drv in ("ev", "evd"):
                    O.scipy.linalg.eigh = lambda M, lower=True, _d=drv: orig(M, lower=lower, driver=_d)
                    f = O.train_dense(L, *a)
                    rel.append(np.abs(f.eigvals - base.eigvals).max() if f.eigvals.size != base.eigvals.size
                               else (np.abs(f.eigvals - base.eigvals) / base.eigvals).max())
            finally:
                O.scipy.linalg.eigh = orig
            f = O.train_streaming(L, *a)
            rel.append((np.abs(f.eigvals - base.eigvals) / base.eigvals).max())
            rel.append(oracle_sq_spread(L, a)[1].max())       # all of the above + 8 draws of Ka + E, ||E|| = eps ||Ka||
            worst = max(worst, *rel)
            out.append(f"| {hx:g} | {hy:g} | {base.stages['r']} / {base.stages['r2']} | {base.eigvals[-1]:.2e} | "
                       + " | ".join(f"{x:.1e}" for x in rel) + " |")
            print(out[-1], file=sys.stderr, flush=True)
    out += ["", f"Reading: eleven of the twelve corners are pinned to better than 1e-6 by every variant.  At hx = 5000, hy = 100 (Ka of rank 32 of",
            "99 at the 1e-10 cut, Sq down to 6e-7) three backward-stable eigensolvers applied to the SAME FP64 matrices move the small",
            f"eigenvalues by up to {worst:.1e} relative: the reference (Eigen's QR-iteration solver) does not define them to 1e-5 there.",
            "Perturbing Ka by a random symmetric E with ‖E‖₂ = ε‖Ka‖₂ -- less than any backward-stable eigensolver is allowed -- moves them",
            "by 1e-4 … 3e-4 as well (last column).  The CUDA path lands 4.3e-4 / 4.5e-4 from the MRRR oracle on the two smallest eigenvalues",
            "(gpurun log of 2026-10-18), i.e. inside that cloud.  tests/nle_testlib.py::oracle_sq_spread recomputes the per-eigenvalue spread",
            "inside the test and `sq_close` widens the 1e-5 bound to 10 × spread (a backward error of 10 ε‖Ka‖₂) only for eigenvalues where",
            "that exceeds 1e-5; every other assertion of that case (rank cuts, output within 1 LSB) and every other test keep the plain 1e-5."]
    return out


if __name__ == "__main__":
    main()
