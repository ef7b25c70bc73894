"""GPU: the device 8-bit BGR<->Lab conversion (csrc/lab.cu) against cv2.cvtColor -- the conversion the reference calls
around the spectral filter (filter.cpp:423,440,463).  Exhaustive: all 2^24 BGR triples forward, all 2^24 Lab triples
backward, byte-exact; then the image-level entry points (trainForEnhancement / enhance with the conversion fused on the
device) against the host-OpenCV route on a README example."""
import cv2
import numpy as np
import pytest

import nonlocal_image_edit_b200 as nb
from nle_testlib import load_case, train_args

pytestmark = pytest.mark.gpu


def _cube():
    v = np.arange(256, dtype=np.uint8)
    a, b, c = np.meshgrid(v, v, v, indexing="ij")
    return np.ascontiguousarray(np.stack([a, b, c], -1).reshape(4096, 4096, 3))


def test_bgr2lab_exhaustive_byte_exact():
    img = _cube()
    assert np.array_equal(nb.bgrToLab(img), cv2.cvtColor(img, cv2.COLOR_BGR2Lab))


def test_lab2bgr_exhaustive_byte_exact():
    img = _cube()
    assert np.array_equal(nb.labToBgr(img), cv2.cvtColor(img, cv2.COLOR_Lab2BGR))


def test_ragged_and_empty_sizes():
    rng = np.random.default_rng(3)
    for shape in [(1, 1, 3), (3, 5, 3), (17, 255, 3), (1, 1025, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(nb.bgrToLab(img), cv2.cvtColor(img, cv2.COLOR_BGR2Lab))
        assert np.array_equal(nb.labToBgr(img), cv2.cvtColor(img, cv2.COLOR_Lab2BGR))
    with pytest.raises(nb.NleError):
        nb.bgrToLab(np.zeros((4, 4), np.uint8))


@pytest.mark.parametrize("name", ["forest", "bird"])
def test_image_level_calls_equal_host_opencv_route(name):
    m, img, gold = load_case(name)
    a = train_args(m)
    f = nb.NLEFilter().trainForEnhancement(img, *a)                   # BGR2Lab on the device
    out = f.enhance(img, m["weights"])                                # BGR2Lab, enhance L, Lab2BGR on the device
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab)
    g = nb.NLEFilter().trainFilter(np.ascontiguousarray(lab[:, :, 0]), *a)
    assert np.array_equal(f.eigvals, g.eigvals)                       # identical L channel -> identical pipeline
    lab[:, :, 0] = g.enhanceLuminance(np.ascontiguousarray(lab[:, :, 0]), m["weights"])
    assert np.array_equal(out, cv2.cvtColor(lab, cv2.COLOR_Lab2BGR))
    d = np.abs(out.astype(int) - gold.astype(int))                    # and the reference's own committed output
    assert d.max() <= 2 and (d <= 1).mean() >= 0.999
