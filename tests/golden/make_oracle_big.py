"""Oracle fixtures for the BASELINE.json configurations whose dense intermediates do not fit in host RAM.

Runs the STREAMING FP64 oracle (oracle/nle_oracle.py::train_streaming -- same mathematics as the dense restatement of
filter.cpp:480-502, pinned to it by tests/test_oracle.py::test_streaming_equals_dense) with the C affinity loop
(oracle/nle_oracle_c.c) and records, per configuration: p, r, r2, k', the eigenvalues Sq, the kept eigenvalues of Ka
and Wa next to the 1e-10 cut, a checksum of the sample indices and of the input luminance, and the oracle's enhanced
L channel as PNG.  tests/test_gpu_oracle_big.py compares the CUDA path with these.

  c3      BASELINE configs[2] (the benchmarked one): bench.workload_images(1024, 1024) luminance, 40x40 samples,
          hx=500 hy=30, T=20, k=50, weights 2 3 4 1
  c3x2    the image bench.py trains at --gpus 2 (weak scaling: bench.workload_images(2048, 1024), same grid and widths as c3).
          Its block of Q is 855 x 855, above the size from which the training path solves eig(Q) with the top-k block
          solver (csrc/eig_topk.cu): the fixture that pins THAT path to the oracle
  c4      BASELINE configs[3]: full-resolution data/rock2.jpg (tests/golden/rock2_input.png), 50x50 samples,
          hx=500 hy=10, T=50, k=100, weights 4 3 4 1
  c5crop  BASELINE configs[4] on its top-left 1024x1024 crop: bench.synth_luminance(4096, 4096)[:1024, :1024],
          50x50 samples, hx=500 hy=30, T=20, k=100, weights 2 3 4 1

Run in the build container (8 host threads):  python tests/golden/make_oracle_big.py c3 c4 c5crop
(c3 ~ 15 min, c3x2 ~ 35 min, c4 ~ 25 min, c5crop ~ 25 min).
"""
import hashlib
import json
import os
import sys
import time

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nle_oracle as O  # noqa: E402
import bench  # noqa: E402


def config(name):
    if name == "c3":
        _, lum = bench.workload_images(1024, 1024)
        return lum, (40, 40, 500.0, 30.0, 20, 50), [2.0, 3.0, 4.0, 1.0]
    if name == "c3x2":
        _, lum = bench.workload_images(2048, 1024)
        return lum, (40, 40, 500.0, 30.0, 20, 50), [2.0, 3.0, 4.0, 1.0]
    if name == "c4":
        img = cv2.imread(os.path.join(HERE, "rock2_input.png"))
        lum = np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_BGR2Lab)[:, :, 0])
        return lum, (50, 50, 500.0, 10.0, 50, 100), [4.0, 3.0, 4.0, 1.0]
    if name == "c5crop":
        lum = np.ascontiguousarray(bench.synth_luminance(4096, 4096)[:1024, :1024])
        return lum, (50, 50, 500.0, 30.0, 20, 100), [2.0, 3.0, 4.0, 1.0]
    raise SystemExit(f"unknown configuration {name}")


def main():
    path = os.path.join(HERE, "oracle_big.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name in sys.argv[1:] or ["c3", "c4", "c5crop"]:
        lum, args, weights = config(name)
        t0 = time.time()
        flt = O.train_streaming(lum.astype(np.float64), *args, block_fn=O.affinity_block_c)
        Lout = O.enhance_luminance(flt, lum, weights)
        st = flt.stages
        Ka_all = np.linalg.eigvalsh(st["Ka"])[::-1]
        Wa_all = np.linalg.eigvalsh(np.tril(st["Wa"]) + np.tril(st["Wa"], -1).T)[::-1]
        r, r2 = int(st["r"]), int(st["r2"])
        cv2.imwrite(os.path.join(HERE, f"{name}_oracle_L.png"), Lout, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        out[name] = dict(
            rows=int(lum.shape[0]), cols=int(lum.shape[1]), args=list(args), weights=weights,
            lum_sha1=hashlib.sha1(lum.tobytes()).hexdigest(),
            p=int(st["p"]), r=r, r2=r2, k=int(flt.eigvals.size),
            Sq=[float(x) for x in flt.eigvals],
            sel_sha1=hashlib.sha1(st["perm"][:st["p"]].astype(np.int32).tobytes()).hexdigest(),
            # the eigenvalues on both sides of the two 1e-10 rank cuts: how far the cut is from a tie
            Ka_cut=[float(x) for x in Ka_all[max(0, r - 2):r + 2]],
            Wa_cut=[float(x) for x in Wa_all[max(0, r2 - 2):r2 + 2]],
            L_sha1=hashlib.sha1(Lout.tobytes()).hexdigest(),
            oracle_seconds=round(time.time() - t0, 1), host_threads=os.cpu_count())
        json.dump(out, open(path, "w"), indent=1)
        print(name, {k: out[name][k] for k in ("p", "r", "r2", "k", "Ka_cut", "Wa_cut", "oracle_seconds")},
              file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
