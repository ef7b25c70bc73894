"""Builds the committed golden fixtures from the read-only reference checkout.

Run ONCE in the build container (needs /root/reference and cv2):
    python tests/golden/make_golden.py

For every README sample row (/root/reference/README.md:74-83) it stores
  <name>_input.png   the reference input image decoded by cv2.imread, re-encoded losslessly
                     (two inputs are JPEGs; storing the decoded pixels removes any dependence on
                     the libjpeg build of the box that later runs the tests),
  <name>_golden.png  the reference's own committed output data/<name>-filtered.{png,bmp},
and writes manifest.json with the CLI parameters of that row.  Nothing here is reference
source code; these are the reference's data files and the parameters its README lists.
"""
import json
import os
import sys

import cv2

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# name, input file, golden file, README.md line, argv[3..] of `enhance`
ROWS = [
    ("flower",        "flower-50.bmp",        "flower-filtered.png",        74, "10 20 100 30 50 30 2 3 4 1"),
    ("bird",          "bird.bmp",             "bird-filtered.png",          75, "10 20 1000 20 10 10 1 5 5 1"),
    ("canyon",        "canyon-dawn-20.bmp",   "canyon-filtered.bmp",        76, "20 10 500 30 40 10 2 7 5 1"),
    ("brickwall",     "brickwall-20.bmp",     "brickwall-filtered.png",     77, "10 20 1000 25 30 50 2 3 3 1"),
    ("conifer",       "conifer-10.bmp",       "conifer-filtered.png",       78, "25 15 800 20 40 100 2 3 5 1"),
    ("forest",        "forest-10.bmp",        "forest-filtered.png",        79, "20 10 5000 30 10 10 4 6 6 1.05"),
    ("snow-mountain", "snow-mountain-15.bmp", "snow-mountain-filtered.png", 80, "10 20 200 30 30 10 3 10 1 1"),
    ("paper",         "paper.jpg",            "paper-filtered.png",         81, "20 20 1000 40 50 20 0.5 1 5 1"),
    ("rock2",         "rock2.jpg",            "rock2-filtered.png",         82, "20 30 500 10 50 50 4 3 4 1"),
    ("red-cherries",  "red-cherries-10.bmp",  "red-cherries-filtered.png",  83, "20 10 400 30 50 20 2 2 2 1"),
]


def main():
    manifest = []
    for name, fin, fgold, line, params in ROWS:
        img = cv2.imread(os.path.join(REF, "data", fin))
        gold = cv2.imread(os.path.join(REF, "data", fgold))
        assert img is not None and gold is not None, name
        assert img.shape == gold.shape, (name, img.shape, gold.shape)
        cv2.imwrite(os.path.join(HERE, f"{name}_input.png"), img, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        cv2.imwrite(os.path.join(HERE, f"{name}_golden.png"), gold, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        a = params.split()
        manifest.append(dict(
            name=name, source_input=f"data/{fin}", source_golden=f"data/{fgold}",
            readme_line=line, rows=int(img.shape[0]), cols=int(img.shape[1]),
            n_row_samples=int(a[0]), n_col_samples=int(a[1]), hx=float(a[2]), hy=float(a[3]),
            n_sinkhorn_iter=int(a[4]), n_eigen_vectors=int(a[5]),
            weights=[float(x) for x in a[6:]]))
        print(name, img.shape, file=sys.stderr)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
