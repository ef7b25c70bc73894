"""Host-side mirror of the reference's include/filter.hpp on top of the C ABI (include/nle_b200.h).

Same names, argument meaning and error behaviour as the reference:
  nle::computeKernel / eigenDecomposition / nystromApproximation / sinkhorn / orthogonalize
  (filter.hpp:20-33) and class nle::NLEFilter (filter.hpp:35-54).
Matrices cross this boundary as NumPy float64 arrays; they are handed to the library in
column-major order (Eigen's default).  Images are BGR uint8 arrays like cv::Mat.

trainForEnhancement / enhance run the 8-bit BGR<->Lab conversion on the device (csrc/lab.cu, byte-exact with
cv::cvtColor; `bgrToLab` / `labToBgr` expose it).  cv2.bilateralFilter, which only the denoise variant uses, stays on
the host where the reference calls OpenCV (filter.cpp:366-371, 535).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NleError, check

EPS = 1e-10


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(a):
    """2-D float64 array laid out column-major (Fortran order); returns (array, pointer)."""
    a = np.asfortranarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
def sampleIndices(rows, cols, nRowSamples, nColSamples, with_rest=True):
    """samplePixels + to1DIndex (filter.cpp:56-80, utils.hpp:11-14): (selected, rest) raster indices."""
    lib = _lib.load()
    p = C.c_int(0)
    check(lib.nle_b200_sample_count(rows, cols, nRowSamples, nColSamples, C.byref(p)))
    sel = np.empty(p.value, dtype=np.int32)
    rest = np.empty(rows * cols - p.value, dtype=np.int32) if with_rest else None
    check(lib.nle_b200_sample_indices(rows, cols, nRowSamples, nColSamples, _ptr(sel),
                                      _ptr(rest) if rest is not None else None))
    return sel, rest


def computeKernel(mat, nRowSamples, nColSamples, hx, hy, with_kab=True):
    """nle::computeKernel (filter.cpp:114-167) -> (P, Ka, Kab); P is the permutation index vector."""
    lib = _lib.load()
    mat = _f64(mat)
    rows, cols = mat.shape
    p = C.c_int(0)
    check(lib.nle_b200_sample_count(rows, cols, nRowSamples, nColSamples, C.byref(p)))
    p = p.value
    n = rows * cols
    perm = np.empty(n, dtype=np.int32)
    Ka = np.empty((p, p), dtype=np.float64, order="F")
    Kab = np.empty((p, n - p), dtype=np.float64, order="F") if with_kab else None
    check(lib.nle_b200_compute_kernel(_ptr(mat), rows, cols, nRowSamples, nColSamples, float(hx), float(hy),
                                      _ptr(perm), _ptr(Ka), _ptr(Kab) if Kab is not None else None))
    return perm, Ka, Kab


def eigenDecomposition(M, eps=EPS):
    """nle::eigenDecomposition (filter.cpp:204-228) -> (U n x r, D r)."""
    lib = _lib.load()
    M, pM = _colmajor(M)
    n = M.shape[0]
    if M.shape != (n, n):
        raise ValueError("M must be square")
    U = np.empty((n, n), dtype=np.float64, order="F")
    D = np.empty(n, dtype=np.float64)
    r = C.c_int(0)
    check(lib.nle_b200_eigen_decomposition(pM, n, float(eps), _ptr(U), _ptr(D), C.byref(r)))
    return np.asfortranarray(U[:, :r.value]), D[:r.value].copy()


def topkEigenDecomposition(M, nLargest, eps=EPS, assume_psd=False, return_products=False):
    """nle::topkEigenDecomposition (filter.cpp:169-200, USE_SPECTRA build) -> (U n x r, D r): the min(nLargest, n-1) eigenpairs
    of largest magnitude in descending order, cut at the first eigenvalue below eps.  assume_psd=True (the reference's only
    caller passes Q, positive semi-definite) selects the block solver the training path uses for eig(Q)."""
    lib = _lib.load()
    M, pM = _colmajor(M)
    n = M.shape[0]
    if M.shape != (n, n):
        raise ValueError("M must be square")
    nev = max(1, min(int(nLargest), n - 1))
    U = np.empty((n, nev), dtype=np.float64, order="F")
    D = np.empty(nev, dtype=np.float64)
    r, prod = C.c_int(0), C.c_int(0)
    check(lib.nle_b200_topk_eigen_decomposition(pM, n, int(nLargest), float(eps), int(bool(assume_psd)), _ptr(U), _ptr(D),
                                                C.byref(r), C.byref(prod)))
    out = (np.asfortranarray(U[:, :r.value]), D[:r.value].copy())
    return out + (prod.value,) if return_products else out


def nystromApproximation(Ka, Kab):
    """nle::nystromApproximation (filter.cpp:257-280) -> (eigvals r, phi N x r)."""
    lib = _lib.load()
    Ka, pKa = _colmajor(Ka)
    Kab, pKab = _colmajor(Kab)
    p, nrest = Ka.shape[0], Kab.shape[1]
    eigvals = np.empty(p, dtype=np.float64)
    phi = np.empty((p + nrest, p), dtype=np.float64, order="F")
    r = C.c_int(0)
    check(lib.nle_b200_nystrom_approximation(pKa, p, pKab, nrest, _ptr(eigvals), _ptr(phi), C.byref(r)))
    return eigvals[:r.value].copy(), np.asfortranarray(phi[:, :r.value])


def sinkhorn(phi, eigvals, maxIter=10):
    """nle::sinkhorn (filter.cpp:230-254) -> (Wa r x r, Wab r x (n-r))."""
    lib = _lib.load()
    phi, pphi = _colmajor(phi)
    eigvals = _f64(eigvals)
    n, r = phi.shape
    Wa = np.empty((r, r), dtype=np.float64, order="F")
    Wab = np.empty((r, n - r), dtype=np.float64, order="F")
    check(lib.nle_b200_sinkhorn(pphi, n, r, _ptr(eigvals), int(maxIter), _ptr(Wa), _ptr(Wab) if n > r else None))
    return Wa, Wab


def orthogonalize(Wa, Wab, nEigVectors=5, eps=EPS):
    """nle::orthogonalize (filter.cpp:282-331) -> (V (p+nrest) x k', S k')."""
    lib = _lib.load()
    Wa, pWa = _colmajor(Wa)
    Wab, pWab = _colmajor(Wab)
    p, nrest = Wa.shape[0], Wab.shape[1]
    V = np.empty((p + nrest, nEigVectors), dtype=np.float64, order="F")
    S = np.empty(nEigVectors, dtype=np.float64)
    k = C.c_int(0)
    check(lib.nle_b200_orthogonalize(pWa, p, pWab if nrest else None, nrest, int(nEigVectors), float(eps),
                                     _ptr(V), _ptr(S), C.byref(k)))
    return np.asfortranarray(V[:, :k.value]), S[:k.value].copy()


def _image_u8(image, what):
    """The reference's images are CV_8UC3 cv::Mat (enhance.cpp:33 imread): anything else is refused, never reinterpreted."""
    img = np.asarray(image)
    if img.dtype != np.uint8:
        raise NleError(-1, f"{what}: expected a uint8 (CV_8U) image, got dtype {img.dtype}")
    return np.ascontiguousarray(img)


def bgrToLab(image):
    """cv::cvtColor(image, COLOR_BGR2Lab) on CV_8UC3 (filter.cpp:423,463), on the device."""
    img = _image_u8(image, "bgrToLab")
    if img.ndim != 3 or img.shape[2] != 3:
        raise NleError(-1, "expected an H x W x 3 uint8 image")
    out = np.empty_like(img)
    check(_lib.load().nle_b200_bgr_to_lab_u8(_ptr(img), img.shape[0] * img.shape[1], _ptr(out)))
    return out


def labToBgr(lab):
    """cv::cvtColor(lab, COLOR_Lab2BGR) on CV_8UC3 (filter.cpp:440), on the device."""
    img = _image_u8(lab, "labToBgr")
    if img.ndim != 3 or img.shape[2] != 3:
        raise NleError(-1, "expected an H x W x 3 uint8 image")
    out = np.empty_like(img)
    check(_lib.load().nle_b200_lab_to_bgr_u8(_ptr(img), img.shape[0] * img.shape[1], _ptr(out)))
    return out


def transformEigenValues(eigvals, weights):
    """transformEigenValues (filter.cpp:334-347)."""
    lib = _lib.load()
    eigvals = _f64(eigvals)
    weights = _f64(weights)
    out = np.empty_like(eigvals)
    check(lib.nle_b200_transform_eigenvalues(_ptr(eigvals), eigvals.size, _ptr(weights), weights.size, _ptr(out)))
    return out


# ---------------------------------------------------------------------------------------------
class NLEFilter:
    """nle::NLEFilter (filter.hpp:35-54).  m_eigvecs / m_eigvals stay resident in HBM behind a handle."""

    def __init__(self):
        self._h = None
        self._lib = _lib.load()

    # -- lifetime --------------------------------------------------------------------------
    def _release(self):
        if getattr(self, "_h", None):
            self._lib.nle_b200_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _require_trained(self):
        if not self._h:
            raise NleError(-1, "filter has not been trained")

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _lab(image):
        return bgrToLab(image)          # cv::cvtColor(COLOR_BGR2Lab) on the device, byte-exact (csrc/lab.cu)

    @staticmethod
    def _bgr(lab):
        return labToBgr(lab)            # cv::cvtColor(COLOR_Lab2BGR) on the device

    def info(self):
        self._require_trained()
        inf = _lib.Info()
        check(self._lib.nle_b200_filter_info(self._h, C.byref(inf)))
        return inf

    @property
    def eigvals(self):
        """m_eigvals (filter.hpp:53)."""
        k = self.info().k
        S = np.empty(k, dtype=np.float64)
        check(self._lib.nle_b200_eigenvalues(self._h, _ptr(S)))
        return S

    @property
    def eigvecs(self):
        """m_eigvecs (filter.hpp:52): N x k, pixel order (after the un-permute of filter.cpp:502)."""
        inf = self.info()
        n = (inf.row1 - inf.row0) * inf.cols
        V = np.empty((n, inf.k), dtype=np.float64, order="F")
        check(self._lib.nle_b200_eigenvectors(self._h, _ptr(V)))
        return V

    def stage(self, which):
        self._require_trained()
        size = C.c_size_t(0)
        rc = self._lib.nle_b200_get_stage(self._h, int(which), None, 0, C.byref(size))
        if rc != 0 and size.value == 0:
            check(rc)
        out = np.empty(size.value, dtype=np.float64)
        check(self._lib.nle_b200_get_stage(self._h, int(which), _ptr(out), out.size, C.byref(size)))
        return out

    # -- trainFilter (filter.cpp:480-502) on a luminance channel -----------------------------
    def trainFilter(self, channel, nRowSamples, nColSamples, hx, hy, nSinkhornIter, nEigenVectors,
                    shard=None):
        """channel: H x W luminance (uint8, or float64 holding integer values as in the reference).

        shard: optional (row0, row1, allreduce) for row-sharded multi-GPU training; `allreduce` is a
        Python callable (dev_ptr:int, count:int, stream:int) -> None summing in place across ranks.
        """
        self._release()
        h = C.c_void_p()
        ch = np.asarray(channel)
        if ch.ndim != 2:
            raise NleError(-1, "channel must be 2-D")
        rows, cols = ch.shape
        if shard is not None:
            row0, row1, allreduce = shard
            ch8 = np.ascontiguousarray(ch, dtype=np.uint8)
            if ch.dtype != np.uint8 and not np.array_equal(ch8, ch):
                raise NleError(-3, "luminance channel must hold integer values in [0,255]")

            def _cb(buf, count, stream, user):
                try:
                    allreduce(buf, count, stream)
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            self._cb = _lib.ALLREDUCE_FN(_cb)
            check(self._lib.nle_b200_train_u8_sharded(_ptr(ch8), rows, cols, int(row0), int(row1), int(nRowSamples),
                                                      int(nColSamples), float(hx), float(hy), int(nSinkhornIter),
                                                      int(nEigenVectors), self._cb, None, C.byref(h)))
        elif ch.dtype == np.uint8:
            ch8 = np.ascontiguousarray(ch)
            check(self._lib.nle_b200_train_u8(_ptr(ch8), rows, cols, int(nRowSamples), int(nColSamples), float(hx),
                                              float(hy), int(nSinkhornIter), int(nEigenVectors), C.byref(h)))
        else:
            chd = _f64(ch)
            check(self._lib.nle_b200_train(_ptr(chd), rows, cols, int(nRowSamples), int(nColSamples), float(hx),
                                           float(hy), int(nSinkhornIter), int(nEigenVectors), C.byref(h)))
        self._h = h
        return self

    # -- public API of the reference class --------------------------------------------------
    def trainForEnhancement(self, image, nRowSamples, nColSamples, hx, hy, nSinkhornIter=10, nEigenVectors=5):
        """filter.cpp:514-519: getLuminanceChannel (:460-469, BGR2Lab on the device) + trainFilter."""
        image = _image_u8(image, "trainForEnhancement")
        if image.ndim != 3 or image.shape[2] != 3:
            raise NleError(-1, "expected an H x W x 3 uint8 image")
        self._release()
        h = C.c_void_p()
        rows, cols = image.shape[:2]
        check(self._lib.nle_b200_train_bgr_u8(_ptr(image), rows, cols, 0, rows, int(nRowSamples), int(nColSamples),
                                              float(hx), float(hy), int(nSinkhornIter), int(nEigenVectors),
                                              C.cast(None, _lib.ALLREDUCE_FN), None, C.byref(h)))
        self._h = h
        return self

    def trainForDenoise(self, image, nRowSamples, nColSamples, hx, hy, nSinkhornIter, nEigenVectors,
                        sigmaColor=10, sigmaSpace=10):
        """filter.cpp:521-538."""
        import cv2
        L = np.ascontiguousarray(self._lab(_image_u8(image, "trainForDenoise"))[:, :, 0])
        den = cv2.bilateralFilter(L, -1, sigmaColor, sigmaSpace, borderType=cv2.BORDER_DEFAULT)
        return self.trainFilter(den, nRowSamples, nColSamples, hx, hy, nSinkhornIter, nEigenVectors)

    def apply(self, channel, transformedEigVals):
        """NLEFilter::apply (filter.cpp:445-458): V diag(fS) V^T channel, float64 in/out."""
        self._require_trained()
        inf = self.info()
        ch = _f64(channel)
        if ch.size != (inf.row1 - inf.row0) * inf.cols:
            raise NleError(-1, "Number of values in channel must match that of training image.")   # :448
        fS = _f64(transformedEigVals)
        if fS.size != inf.k:
            raise NleError(-1, "transformed eigenvalues must have one entry per eigenvector")
        out = np.empty_like(ch)
        check(self._lib.nle_b200_apply(self._h, _ptr(ch), ch.size, _ptr(fS), _ptr(out)))
        return out

    def enhanceLuminance(self, lum_u8, weights):
        """filter.cpp:426-436 on the 8-bit L channel, fused on the device."""
        self._require_trained()
        inf = self.info()
        lum = _image_u8(lum_u8, "enhanceLuminance")
        if lum.size != (inf.row1 - inf.row0) * inf.cols:
            raise NleError(-1, "Cannot apply filter on image with different size from the image filter was trained on.")
        w = _f64(weights)
        out = np.empty_like(lum)
        check(self._lib.nle_b200_enhance_luminance_u8(self._h, _ptr(lum), _ptr(w), w.size, _ptr(out)))
        return out

    def enhance(self, image, weights):
        """NLEFilter::enhance (filter.cpp:412-443)."""
        image = _image_u8(image, "enhance")
        if image.ndim != 3 or image.shape[2] != 3:
            raise NleError(-1, "Can only enhance RGB image.")                                    # :415
        self._require_trained()
        inf = self.info()
        if image.shape[0] * image.shape[1] != inf.rows * inf.cols:
            raise NleError(-1, "Cannot apply filter on image with different size from the image filter was trained on.")  # :419
        if (inf.row0, inf.row1) != (0, inf.rows):
            raise NleError(-1, "enhance() needs the whole image on this filter; use enhanceLuminance on a row slab")
        w = _f64(weights)
        out = np.empty_like(image)
        check(self._lib.nle_b200_enhance_bgr_u8(self._h, _ptr(image), image.shape[0], image.shape[1], image.shape[2],
                                                _ptr(w), w.size, _ptr(out)))                      # :422-440 on the device
        return out

    def denoise(self, image, k, sigmaColor=10, sigmaSpace=10):
        """NLEFilter::denoise (filter.cpp:349-410) without the imshow side effects."""
        import cv2
        image = _image_u8(image, "denoise")
        if image.ndim != 3 or image.shape[2] != 3:
            raise NleError(-1, "Can only enchance RGB image.")                                   # :352 (sic)
        self._require_trained()
        inf = self.info()
        if image.shape[0] * image.shape[1] != inf.rows * inf.cols:
            raise NleError(-1, "Cannot apply filter on image with different size from the image filter was trained on.")  # :356
        if (inf.row0, inf.row1) != (0, inf.rows):
            raise NleError(-1, "denoise() needs the whole image on this filter (it was trained on a row slab)")
        lab = self._lab(image)
        Y = cv2.bilateralFilter(np.ascontiguousarray(lab[:, :, 0]), -1, sigmaColor, sigmaSpace,
                                borderType=cv2.BORDER_DEFAULT)                                   # :371
        out = np.empty_like(lab)
        out[:, :, 0] = Y                                # channel 0 is not filtered (:387 is commented out)
        for c in (1, 2):                                                                         # :388-389
            src = np.ascontiguousarray(lab[:, :, c])
            dst = np.empty_like(src)
            check(self._lib.nle_b200_denoise_channel_u8(self._h, _ptr(src), src.shape[0], src.shape[1], float(k), _ptr(dst)))
            out[:, :, c] = dst
        return self._bgr(out)
