"""ctypes binding of libnle_b200.so (C ABI: include/nle_b200.h).  Loads the in-tree shared library;
there is no CPU fallback -- if the library is missing it is built with nvcc, and every compute call
fails loudly when no CUDA device is present."""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnle_b200.so")

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)


class Info(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("row0", C.c_int), ("row1", C.c_int),
                ("p", C.c_int), ("r", C.c_int), ("r2", C.c_int), ("k", C.c_int),
                ("n_row_samples_eff", C.c_int), ("n_col_samples_eff", C.c_int),
                ("eig_sweeps", C.c_int * 3), ("eig_fallbacks", C.c_int), ("topk_products", C.c_int)]


# every symbol include/nle_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_I32 = C.POINTER(C.c_int32)
_U8 = C.POINTER(C.c_uint8)
SYMBOLS = {
    "nle_b200_last_error": (C.c_char_p, []),
    "nle_b200_version": (C.c_int, []),
    "nle_b200_device_count": (C.c_int, []),
    "nle_b200_sample_count": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _I]),
    "nle_b200_sample_indices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nle_b200_compute_kernel": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _P, _P, _P]),
    "nle_b200_eigen_decomposition": (C.c_int, [_P, C.c_int, C.c_double, _P, _P, _I]),
    "nle_b200_topk_eigen_decomposition": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int, _P, _P, _I, _I]),
    "nle_b200_nystrom_approximation": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _I]),
    "nle_b200_sinkhorn": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "nle_b200_orthogonalize": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_double, _P, _P, _I]),
    "nle_b200_transform_eigenvalues": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "nle_b200_train": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(_P)]),
    "nle_b200_train_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(_P)]),
    "nle_b200_train_u8_sharded": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, ALLREDUCE_FN, _P, C.POINTER(_P)]),
    "nle_b200_train_u8_dev": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, ALLREDUCE_FN, _P, C.POINTER(_P)]),
    "nle_b200_filter_info": (C.c_int, [_P, C.POINTER(Info)]),
    "nle_b200_eigenvalues": (C.c_int, [_P, _P]),
    "nle_b200_eigenvectors": (C.c_int, [_P, _P]),
    "nle_b200_apply": (C.c_int, [_P, _P, C.c_longlong, _P, _P]),
    "nle_b200_enhance_luminance_u8": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "nle_b200_enhance_luminance_u8_dev": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "nle_b200_denoise_channel_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_double, _P]),
    "nle_b200_bgr_to_lab_u8": (C.c_int, [_P, C.c_longlong, _P]),
    "nle_b200_lab_to_bgr_u8": (C.c_int, [_P, C.c_longlong, _P]),
    "nle_b200_train_bgr_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, ALLREDUCE_FN, _P, C.POINTER(_P)]),
    "nle_b200_enhance_bgr_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "nle_b200_get_stage": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "nle_b200_set_keep_stages": (None, [C.c_int]),
    "nle_b200_launch_count": (C.c_longlong, [C.c_int]),
    "nle_b200_free": (None, [_P]),
    "nle_b200_release_cache": (None, []),
    "nle_b200_comm_unique_id": (C.c_int, [_P]),
    "nle_b200_comm_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "nle_b200_comm_allreduce": (C.c_int, [_P, C.c_size_t, _P, _P]),
    "nle_b200_comm_info": (C.c_char_p, [_P, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]),
    "nle_b200_comm_destroy": (None, [_P]),
    "nle_b200_measured_peaks": (C.c_int, [_P, C.c_int]),
    "nle_b200_fp64_fma_peak_tflops": (C.c_double, []),
    "nle_b200_fp64_dmma_peak_tflops": (C.c_double, []),
}

_lib = None


def load():
    """Returns the loaded CDLL, building it in-tree first if necessary."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the ABI symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class NleError(RuntimeError):
    """Mirrors the std::runtime_error the reference throws (filter.cpp:118,415,419,448)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def check(code):
    if code != 0:
        msg = load().nle_b200_last_error()
        raise NleError(code, (msg or b"").decode("utf-8", "replace"))
