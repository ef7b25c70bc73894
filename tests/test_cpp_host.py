"""The C++ host mirror (include/nle_b200.hpp: class NLEFilter and the free functions with the reference's names over
plain buffers) and its test program tests/cpp/host_mirror_test.cpp: the C ABI driven from C++ without Python, OpenCV or
Eigen.  CPU: it builds, links every symbol it uses, and fails loudly without a device.  GPU: a README example through
the C++ path equals the Python path byte for byte."""
import os
import subprocess

import numpy as np
import pytest

from nle_testlib import load_case, train_args
from nonlocal_image_edit_b200 import build as nbuild


@pytest.fixture(scope="module")
def exe():
    return nbuild.build_host_test()


def _args(m, rows, cols, inp, out):
    a = train_args(m)
    return [inp, str(rows), str(cols), str(a[0]), str(a[1]), repr(float(a[2])), repr(float(a[3])), str(a[4]), str(a[5]), out] + \
           [repr(float(w)) for w in m["weights"]]


def test_builds_and_prints_usage(exe):
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 2 and "usage" in res.stderr


def test_fails_loudly_without_a_device(exe, tmp_path):
    import nonlocal_image_edit_b200 as nb
    if nb.load().nle_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    m, img, _ = load_case("forest")
    inp = str(tmp_path / "in.raw")
    img.tofile(inp)
    res = subprocess.run([exe] + _args(m, img.shape[0], img.shape[1], inp, str(tmp_path / "out.raw")), capture_output=True, text=True)
    assert res.returncode == 1 and "FAIL" in res.stderr       # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["forest", "brickwall"])
def test_cpp_path_equals_python_path(exe, tmp_path, name):
    import nonlocal_image_edit_b200 as nb
    m, img, gold = load_case(name)
    inp, outp = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    np.ascontiguousarray(img).tofile(inp)
    res = subprocess.run([exe] + _args(m, img.shape[0], img.shape[1], inp, outp), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert res.stdout.startswith("ok ")
    out = np.fromfile(outp, dtype=np.uint8).reshape(img.shape)
    f = nb.NLEFilter().trainForEnhancement(img, *train_args(m))
    assert np.array_equal(out, f.enhance(img, m["weights"]))
    inf = f.info()
    assert f"p={inf.p} r={inf.r} r2={inf.r2} k={inf.k}" in res.stdout
    d = np.abs(out.astype(int) - gold.astype(int))
    assert d.max() <= 2 and (d <= 1).mean() >= 0.999
