"""nonlocal_image_edit_b200 -- B200-native Nystrom spectral filter (hot path of
lightalchemist/nonlocal-image-edit) behind the reference's own operator interface.

Host side mirrors include/filter.hpp of the reference: class NLEFilter with trainForEnhancement /
trainForDenoise / enhance / denoise, and the free functions computeKernel, eigenDecomposition,
nystromApproximation, sinkhorn, orthogonalize.  All numerics run in libnle_b200.so (hand-written
CUDA for sm_100a), including the 8-bit BGR<->Lab conversion around enhance; Python only does what the
reference does with OpenCV on the host for I/O (imread) and for the denoise variant (bilateralFilter)."""
from .filter import (NLEFilter, bgrToLab, computeKernel, eigenDecomposition, labToBgr,  # noqa: F401
                     nystromApproximation, orthogonalize, sampleIndices, sinkhorn, topkEigenDecomposition,
                     transformEigenValues)
from ._lib import NleError, load  # noqa: F401

EPS = 1e-10  # include/filter.hpp:14
