"""Developer diagnostic: A/B two builds of a stage selected by an environment variable, e.g.
     python scripts/gpu_ab_check.py NLE_B200_GRAM=pixel        (pixel-axis SYRK vs the default cell Gram)
     python scripts/gpu_ab_check.py NLE_B200_SINKHORN=rows     (per-row Sinkhorn kernels vs the level-table GEMMs)
Prints the relative differences of the stage outputs and the per-stage device times."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(96, 128, 8, 10, 40.0, 25.0, 6, 8), (200, 333, 20, 10, 300.0, 30.0, 5, 10), (512, 512, 40, 40, 500.0, 30.0, 4, 20),
         (300, 260, 50, 50, 200.0, 10.0, 3, 20), (64, 48, 3, 1, 20.0, 15.0, 3, 2), (1024, 1024, 40, 40, 500.0, 30.0, 20, 50)]

def child(out):
    import nonlocal_image_edit_b200 as nb
    from bench import synth_luminance
    nb.load().nle_b200_set_keep_stages(1)
    res = {}
    for i, (h, w, a, b, hx, hy, T, k) in enumerate(CASES):
        L = synth_luminance(h, w, seed=7 + i)
        for rep in range(2):   # second run: warm timings
            f = nb.NLEFilter().trainFilter(L, a, b, hx, hy, T, k)
        for st, nm in ((2, "rvec"), (3, "c"), (7, "G")):
            res[f"{nm}{i}"] = f.stage(st)
        res[f"S{i}"] = f.eigvals
        res[f"t{i}"] = f.stage(8)
        res[f"out{i}"] = f.enhanceLuminance(L, [2.0, 3.0, 4.0, 1.0])
    np.savez(out, **res)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2]); sys.exit(0)
    var, val = sys.argv[1].split("=")
    outs = {}
    for mode in ("alt", "default"):
        env = dict(os.environ)
        env.pop(var, None)
        if mode == "alt": env[var] = val
        o = f"/tmp/ab_{mode}.npz"
        subprocess.run([sys.executable, __file__, "--child", o], env=env, check=True)
        outs[mode] = np.load(o)
    rel = lambda a, b: (np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)) if a.shape == b.shape else float('nan')
    for i, c in enumerate(CASES):
        A, B = outs["alt"], outs["default"]
        d = np.abs(A[f"out{i}"].astype(int) - B[f"out{i}"].astype(int))
        print(c, " ".join(f"{nm} {rel(A[f'{nm}{i}'], B[f'{nm}{i}']):.2e}" for nm in ("rvec", "c", "G", "S")),
              f"out max {d.max()} differing {(d != 0).mean():.2e}")
        print("    stage ms [setup eigKa sinkhorn gram small ext total gramk]  alt", np.round(A[f"t{i}"], 2), " default", np.round(B[f"t{i}"], 2))
