"""CPU: the literal drop-in (integration/filter_b200.cpp, the replacement for the reference's src/filter.cpp) type-checks
against the reference's OWN, unchanged include/filter.hpp and include/utils.hpp.  Eigen and the OpenCV C++ SDK are not in
this image, so the third-party headers are the minimal declaration-only stand-ins under tests/cpp/stubs/ and the check is
g++ -fsyntax-only (nothing is linked or run).  /root/reference does not exist on the GPU box: skipped there."""
import os
import re
import subprocess

import pytest

from nle_testlib import ROOT

REF_INC = "/root/reference/include"
SHIM = os.path.join(ROOT, "integration", "filter_b200.cpp")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_INC, "filter.hpp")), reason="reference headers not mounted")
def test_shim_type_checks_against_the_unchanged_reference_header():
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-function",
           "-I", os.path.join(ROOT, "tests", "cpp", "stubs"), "-I", REF_INC, "-I", os.path.join(ROOT, "include"), SHIM]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_shim_defines_every_function_filter_hpp_declares_and_uses_only_declared_c_symbols():
    src = open(SHIM).read()
    for name in ("computeKernel", "eigenDecomposition", "nystromApproximation", "sinkhorn", "orthogonalize",
                 "NLEFilter::trainForEnhancement", "NLEFilter::trainForDenoise", "NLEFilter::enhance", "NLEFilter::denoise",
                 "NLEFilter::apply", "NLEFilter::trainFilter"):
        assert re.search(r"\b" + re.escape(name) + r"\s*\(", src), name
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "nle_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(nle_b200_[a-z0-9_]+)\s*\(", hdr))
    used = set(re.findall(r"\b(nle_b200_[a-z0-9_]+)\s*\(", src))
    assert used and used <= declared, used - declared
    # the reference's runtime_error strings that the shim itself raises (the others come from the C ABI)
    assert "Can only enchance RGB image." in src
    assert "Cannot apply filter on image with different size from the image filter was trained on." in src
