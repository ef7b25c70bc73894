"""CPU tests (-m "not gpu"): the C-ABI library builds for sm_100a, loads, exports every declared symbol;
host-side logic (grid arithmetic, error mirroring, sharding helper) works without a GPU; the product
never routes through the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nle_testlib import ROOT
from oracle import nle_oracle as O


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nle_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nle_b200_[a-z0-9_]+)\s*\(", hdr)) - {"nle_b200_allreduce_fn"})


def test_library_builds_and_exports_every_declared_symbol():
    from nonlocal_image_edit_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nle_b200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"
    assert lib.nle_b200_version() >= 100


def test_sm100a_code_is_embedded():
    from nonlocal_image_edit_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sample_count_is_host_arithmetic(nb):
    lib = nb.load()
    p = C.c_int(0)
    for shape in [(736, 491, 20, 10), (100, 100, 40, 40), (1024, 1024, 40, 40), (584, 876, 50, 50), (4096, 4096, 50, 50)]:
        assert lib.nle_b200_sample_count(*shape, C.byref(p)) == 0
        sel, _ = O.sample_pixels(*shape) if shape[0] * shape[1] < 2_000_000 else (None, None)
        expect = O.sample_axis(shape[0], shape[2]).size * O.sample_axis(shape[1], shape[3]).size
        assert p.value == expect
        if sel is not None:
            assert sel.size == expect


def test_reference_error_message_for_too_many_samples(nb):
    lib = nb.load()
    p = C.c_int(0)
    rc = lib.nle_b200_sample_count(10, 10, 11, 2, C.byref(p))
    assert rc == -1
    assert lib.nle_b200_last_error().decode() == "Number of samples per row and col must be <= that of image."   # filter.cpp:118


def test_transform_eigenvalues_matches_oracle(nb):
    S = np.array([1.0003, 0.91, 0.5, 1e-6])
    for w in ([4, 6, 6, 1.05], [2, 3, 4, 1], [0.5, 1, 5, 1, 0.9], [1.0]):
        assert np.allclose(nb.transformEigenValues(S, w), O.transform_eigenvalues(S, w), rtol=1e-15, atol=0)


def test_compute_calls_fail_loudly_without_a_gpu(nb):
    lib = nb.load()
    if lib.nle_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(nb.NleError) as e:
        nb.NLEFilter().trainFilter(np.zeros((8, 8), np.uint8), 2, 2, 10.0, 10.0, 2, 2)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(nb.NleError):
        nb.eigenDecomposition(np.eye(3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nonlocal_image_edit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("no oracle", ""), f"{f} mentions the oracle"


def test_row_slab_partition():
    from nonlocal_image_edit_b200.sharding import row_slab
    for rows, world in [(1024, 8), (491, 4), (7, 7), (10, 3)]:
        slabs = [row_slab(rows, r, world) for r in range(world)]
        assert slabs[0][0] == 0 and slabs[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        sizes = [b - a for a, b in slabs]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_slab(4, 0, 8)
