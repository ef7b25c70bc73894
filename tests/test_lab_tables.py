"""CPU: the committed integer tables of the 8-bit BGR<->Lab conversion (csrc/lab_tables.inc) are the ones
scripts/make_lab_tables.py generates, and its NumPy restatement of OpenCV's fixed-point pipeline reproduces cv2.cvtColor
on a random sample of both cubes (the script itself checks all 2^24 + 2^24 triples before it writes the file; the GPU
test tests/test_gpu_lab.py repeats the exhaustive check through the kernels)."""
import importlib.util
import os
import re

import cv2
import numpy as np

from nle_testlib import ROOT


def _script():
    spec = importlib.util.spec_from_file_location("make_lab_tables", os.path.join(ROOT, "scripts", "make_lab_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _committed():
    txt = open(os.path.join(ROOT, "nonlocal_image_edit_b200", "csrc", "lab_tables.inc")).read()
    out = {}
    for name, body in re.findall(r"NLE_LAB_TAB [a-z ]+ (\w+)\[\d+\] = \{([^}]*)\}", txt):
        out[name] = np.array([int(v) for v in body.replace("\n", " ").split(",") if v.strip()], dtype=np.int64)
    return out


def test_committed_tables_equal_the_generator():
    m = _script()
    gamma, cbrt, ytab, fytab, invgamma, cf, ci = m.tables()
    c = _committed()
    assert np.array_equal(c["kLabGammaTab"], gamma) and np.array_equal(c["kLabCbrtTab"], cbrt)
    assert np.array_equal(c["kLabYTab"], ytab) and np.array_equal(c["kLabFyTab"], fytab)
    assert np.array_equal(c["kLabInvGammaTab"], invgamma)
    assert list(c["kLabFwdCoef"]) == cf and list(c["kLabInvCoef"]) == ci


def test_integer_pipeline_reproduces_cv2_on_a_random_sample():
    m = _script()
    T = m.tables()
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (1 << 20, 1, 3), dtype=np.uint8)
    img[:256, 0, :] = np.arange(256, dtype=np.uint8)[:, None]          # the gray axis, where L alone matters
    assert np.array_equal(m.bgr2lab(img, T), cv2.cvtColor(img, cv2.COLOR_BGR2Lab))
    assert np.array_equal(m.lab2bgr(img, T), cv2.cvtColor(img, cv2.COLOR_Lab2BGR))
