"""Shared helpers for the test-suite (fixtures on disk, small synthetic images)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def oracle_stages():
    with open(os.path.join(GOLDEN, "oracle_stages.json")) as f:
        return json.load(f)


def load_case(name):
    import cv2
    m = {x["name"]: x for x in manifest()}[name]
    img = cv2.imread(os.path.join(GOLDEN, f"{name}_input.png"))
    gold = cv2.imread(os.path.join(GOLDEN, f"{name}_golden.png"))
    return m, img, gold


def train_args(m):
    return (m["n_row_samples"], m["n_col_samples"], m["hx"], m["hy"], m["n_sinkhorn_iter"], m["n_eigen_vectors"])


def synth_lum(rows, cols, seed=7):
    """Small structured 8-bit luminance (smooth waves + blocks + mild noise) for fast parity cases."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    img = 120 + 50 * np.sin(x / 9.0) * np.cos(y / 7.0) + 30 * ((x // 16 + y // 12) % 2) + 6 * rng.standard_normal((rows, cols))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


