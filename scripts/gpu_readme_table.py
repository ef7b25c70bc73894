"""Developer diagnostic / evidence: every README example (BASELINE.json configs[0..1]) through the image-level API
(BGR in, BGR out, colour conversion on the device): wall time of train+enhance with warm caches, MP/s, ranks, and the
difference to the reference's own committed output data/*-filtered.png.  Writes a Markdown table to stdout."""
import json, os, sys, time
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb
G = os.path.join(ROOT, "tests/golden")
print("| image | H x W | p | r | r2 | k | T | train+enhance ms | MP/s | max abs diff vs golden | within 1 LSB | identical bytes |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for m in json.load(open(f"{G}/manifest.json")):
    img = cv2.imread(f"{G}/{m['name']}_input.png"); gold = cv2.imread(f"{G}/{m['name']}_golden.png")
    a = (m["n_row_samples"], m["n_col_samples"], m["hx"], m["hy"], m["n_sinkhorn_iter"], m["n_eigen_vectors"])
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        f = nb.NLEFilter().trainForEnhancement(img, *a)
        out = f.enhance(img, m["weights"])
        best = min(best, time.perf_counter() - t0)
    inf = f.info()
    d = np.abs(out.astype(int) - gold.astype(int))
    print(f"| {m['name']} | {img.shape[0]} x {img.shape[1]} | {inf.p} | {inf.r} | {inf.r2} | {inf.k} | {a[4]} | {best*1e3:.1f} | "
          f"{img.shape[0]*img.shape[1]/1e6/best:.2f} | {d.max()} | {(d<=1).mean()*100:.3f} % | {(d==0).mean()*100:.2f} % |", flush=True)
