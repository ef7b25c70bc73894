"""Records stage intermediates of the FP64 oracle on the ten README rows (G5 in SURVEY.md 8c):
p, r, r2, k', the eigenvalues Sq, a checksum of the sample indices and of the oracle's L output.
Run in the build container:  python tests/golden/make_oracle_stages.py   (about 3 minutes)."""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import nle_oracle as O  # noqa: E402


def main():
    man = json.load(open(os.path.join(HERE, "manifest.json")))
    out = {}
    for m in man:
        img = cv2.imread(os.path.join(HERE, f"{m['name']}_input.png"))
        gold = cv2.imread(os.path.join(HERE, f"{m['name']}_golden.png"))
        flt = O.train_for_enhancement(img, m["n_row_samples"], m["n_col_samples"], m["hx"], m["hy"],
                                      m["n_sinkhorn_iter"], m["n_eigen_vectors"])
        lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
        Lout = O.enhance_luminance(flt, np.ascontiguousarray(lab[:, :, 0]), m["weights"])
        cv2.imwrite(os.path.join(HERE, f"{m['name']}_oracle_L.png"), Lout, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        res = O.enhance(flt, img, m["weights"])
        d = np.abs(res.astype(int) - gold.astype(int))
        st = flt.stages
        out[m["name"]] = dict(p=int(st["p"]), r=int(st["r"]), r2=int(st["r2"]), k=int(flt.eigvals.size),
                              Sq=[float(x) for x in flt.eigvals],
                              sel_sha1=hashlib.sha1(st["perm"][:st["p"]].astype(np.int32).tobytes()).hexdigest(),
                              golden_max_diff=int(d.max()), golden_le1=float((d <= 1).mean()))
        print(m["name"], out[m["name"]]["r"], out[m["name"]]["r2"], out[m["name"]]["golden_max_diff"], file=sys.stderr)
    json.dump(out, open(os.path.join(HERE, "oracle_stages.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
