"""Developer diagnostic: cell-contracted Gram (default) against the pixel-axis SYRK (NLE_B200_GRAM=pixel)."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(96, 128, 8, 10, 40.0, 25.0, 6, 8), (200, 333, 20, 10, 300.0, 30.0, 5, 10), (512, 512, 40, 40, 500.0, 30.0, 4, 20),
         (300, 260, 50, 50, 200.0, 10.0, 3, 20), (64, 48, 3, 1, 20.0, 15.0, 3, 2)]

def child(out):
    import nonlocal_image_edit_b200 as nb
    from bench import synth_luminance
    nb.load().nle_b200_set_keep_stages(1)
    res = {}
    for i, (h, w, a, b, hx, hy, T, k) in enumerate(CASES):
        L = synth_luminance(h, w, seed=7 + i)
        f = nb.NLEFilter().trainFilter(L, a, b, hx, hy, T, k)
        res[f"G{i}"] = f.stage(7)
        res[f"S{i}"] = f.eigvals
        res[f"t{i}"] = f.stage(8)
    np.savez(out, **res)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1]); sys.exit(0)
    outs = {}
    for mode in ("pixel", "cells"):
        env = dict(os.environ)
        if mode == "pixel": env["NLE_B200_GRAM"] = "pixel"
        else: env.pop("NLE_B200_GRAM", None)
        o = f"/tmp/gram_{mode}.npz"
        subprocess.run([sys.executable, __file__, o], env=env, check=True)
        outs[mode] = np.load(o)
    for i, c in enumerate(CASES):
        a, b = outs["pixel"][f"G{i}"], outs["cells"][f"G{i}"]
        sa, sb = outs["pixel"][f"S{i}"], outs["cells"][f"S{i}"]
        srel = np.abs(sa - sb).max() / np.abs(sa).max() if sa.size == sb.size else float('nan')
        print(c, "G rel diff %.3e  (|G| max %.3e)  S rel diff %.3e  gram ms pixel %.3f cells %.3f" % (
            np.abs(a - b).max() / np.abs(a).max(), np.abs(a).max(), srel, outs["pixel"][f"t{i}"][7], outs["cells"][f"t{i}"][7]))
