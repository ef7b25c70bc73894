"""Developer check + timing of the tridiagonalisation variants (eig_dc.cu):
   gridsync (default) | resident (NLE_B200_TRD) | resident:sys (volatile cells) | dyn:<cols> (NLE_B200_TRD_DYN: trailing
   columns dealt to ceil(m / cols) CTAs only) | cluster[:S] (NLE_B200_TRD=cluster: clusters of S CTAs share the polling) | grid:<G> (NLE_B200_TRD_GRID: at most G CTAs); join with '+', e.g. resident+grid:64.
   All of them must be bit-identical to gridsync.
   The output of the round-1 runs is condensed in profiles/r1l_trd_phases.md.

  timeout 40 python scripts/gpu_trd_resident.py [n ...] [resident dyn:8 dyn:16 grid:64 resident+grid:74 ...]      # stderr: per-phase ms (NLE_B200_EIG_PROF) and, for the
                                                             # resident kernels, cycles per step by phase (NLE_B200_TRD_PROF)
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb

os.environ["NLE_B200_EIG_STRICT"] = "1"
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1600, 300, 1041, 612, 3, 4, 5, 33, 149, 1800]
MODES = ["gridsync"] + ([a for a in sys.argv[1:] if not a.isdigit()] or ["resident"])
QUICK = os.environ.get("TRD_QUICK") == "1"


def say(msg):
    print(msg, file=sys.stderr, flush=True)


def run(A, mode, reps=3):
    for k in ("NLE_B200_TRD", "NLE_B200_TRD_LL", "NLE_B200_TRD_DYN", "NLE_B200_TRD_GRID", "NLE_B200_TRD_CLUSTER"):
        os.environ.pop(k, None)
    # a mode is '+'-joined: gridsync | resident | resident:sys | cluster[:S] | dyn:<cols per CTA> | grid:<max CTAs>
    for part in mode.split("+"):
        if part.startswith("dyn:"):
            os.environ["NLE_B200_TRD_DYN"] = part[4:]
        elif part.startswith("grid:"):
            os.environ["NLE_B200_TRD_GRID"] = part[5:]
        elif part.startswith("cluster"):
            os.environ["NLE_B200_TRD"] = "cluster"
            if ":" in part:
                os.environ["NLE_B200_TRD_CLUSTER"] = part.split(":")[1]
        elif part.startswith("resident"):
            os.environ["NLE_B200_TRD"] = "resident"
            if part.endswith(":sys"):
                os.environ["NLE_B200_TRD_LL"] = "sys"       # volatile (sys-scope) cells instead of relaxed.gpu
    os.environ.pop("NLE_B200_EIG_PROF", None)
    os.environ.pop("NLE_B200_TRD_PROF", None)
    out = nb.eigenDecomposition(A, eps=-1e300)                    # warm-up (function attributes, pool)
    os.environ["NLE_B200_EIG_PROF"] = "1"
    for _ in range(reps):
        out = nb.eigenDecomposition(A, eps=-1e300)
    os.environ.pop("NLE_B200_EIG_PROF", None)
    if "resident" in mode or "cluster" in mode:
        os.environ["NLE_B200_TRD_PROF"] = "1"
        nb.eigenDecomposition(A, eps=-1e300)
        os.environ.pop("NLE_B200_TRD_PROF", None)
    return out


def matrices(n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, max(3, n // 2)))
    yield "psd", B @ B.T / n + 1e-3 * np.eye(n)
    if n <= 200 and not QUICK:
        yield "identity", np.eye(n)
        yield "zeros", np.zeros((n, n))
        yield "ones", np.ones((n, n))
        yield "diag", np.diag(np.arange(1.0, n + 1.0))


for n in sizes:
    for name, A in matrices(n):
        ref = None
        for mode in MODES:
            say(f"## n={n} {name} mode={mode}")
            try:
                U, D = run(A, mode, reps=3 if name == "psd" else 1)
            except Exception as ex:
                say(f"   FAILED: {ex}")
                continue
            if ref is None:
                ref = (U, D)
                continue
            scale = max(1.0, np.abs(ref[1]).max())
            derr = np.abs(D - ref[1]).max() / scale
            orth = np.abs(U.T @ U - np.eye(n)).max()
            resid = np.abs(A @ U - U * D).max() / scale
            same = np.array_equal(D, ref[1]) and np.array_equal(U, ref[0])
            ok = derr < 1e-13 * max(1, n / 50) and orth < 1e-11 and resid < 1e-12 * max(1, n / 50)
            say(f"   {mode}: bit-identical={same} |D-D0|/max={derr:.2e} orth={orth:.2e} resid={resid:.2e} {'ok' if ok else 'BAD'}")
