"""ctypes binding of librcv_imgproc.so (include/rcv_imgproc.h).

This is the Python twin of the Rust `mod sys { extern "C" { ... } }` block a RustCV
maintainer would add, in the idiom of
rustcv-camera/src/backend/macos/mod.rs:42-80.  There is no CPU fallback: if the
library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCV_IMGPROC_LIB") or os.path.join(_HERE, "librcv_imgproc.so")

RCV_OK = 0
RCV_ERR_ARG, RCV_ERR_SIZE, RCV_ERR_DEPTH, RCV_ERR_CUDA = -1, -2, -3, -4
RCV_ERR_UNSUPPORTED, RCV_ERR_NOT_INIT, RCV_ERR_NOMEM, RCV_ERR_NCCL = -5, -6, -7, -8
RCV_U8, RCV_F32 = 0, 1
RCV_HOST, RCV_DEVICE, RCV_HOST_PINNED = 0, 1, 2

COLOR_YUYV2BGR, COLOR_UYVY2BGR, COLOR_BGRA2BGR, COLOR_RGB2BGR = 0, 1, 2, 3
COLOR_BGR2RGB = 3
COLOR_BGR2GRAY, COLOR_BGR2XRGB32, COLOR_YUYV2GRAY = 4, 5, 6


class RcvMat(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("step", C.c_size_t),
        ("channels", C.c_uint8),
        ("depth", C.c_uint8),
        ("loc", C.c_uint8),
        ("reserved", C.c_uint8),
        ("device", C.c_int32),
    ]


class RcvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rcv error {code}: {msg}")
        self.code = code


_P = C.POINTER
_MatP = _P(RcvMat)

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "rcv_init": [C.c_int],
    "rcv_init_multi": [C.c_int32],
    "rcv_shutdown": [],
    "rcv_device_count": [_P(C.c_int)],
    "rcv_set_blocking": [C.c_int],
    "rcv_sync": [C.c_int],
    "rcv_get_stream": [C.c_int, _P(C.c_void_p)],
    "rcv_launch_count": [_P(C.c_uint64)],
    "rcv_last_error": [],
    "rcv_version": [],
    "rcv_mat_alloc_device": [_MatP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32],
    "rcv_mat_free_device": [_MatP],
    "rcv_mat_alloc_device_batch": [_MatP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32],
    "rcv_mat_free_device_batch": [_MatP, C.c_int32],
    "rcv_mat_upload": [_MatP, _MatP],
    "rcv_mat_download": [_MatP, _MatP],
    "rcv_pinned_alloc": [_P(C.c_void_p), C.c_size_t],
    "rcv_pinned_alloc_on": [C.c_int32, _P(C.c_void_p), C.c_size_t],
    "rcv_pinned_free": [C.c_void_p],
    "rcv_host_register": [C.c_void_p, C.c_size_t],
    "rcv_host_unregister": [C.c_void_p],
    "rcv_cvt_color": [_MatP, _MatP, C.c_int32],
    "rcv_yuyv_to_bgr": [_MatP, _MatP],
    "rcv_yuyv_to_bgr_packed": [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t],
    "rcv_bgra_to_bgr_packed": [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t],
    "rcv_nv12_to_bgr": [_MatP, _MatP, _MatP],
    "rcv_mjpeg_info": [C.c_void_p, C.c_size_t, _P(C.c_int32), _P(C.c_int32)],
    "rcv_mjpeg_to_bgr": [C.c_void_p, C.c_size_t, _MatP],
    "rcv_convert_to": [_MatP, _MatP, C.c_double, C.c_double],
    "rcv_gaussian_blur": [_MatP, _MatP, C.c_int32, C.c_int32, C.c_double, C.c_double],
    "rcv_sep_filter2d": [_MatP, _MatP, _P(C.c_float), C.c_int32, _P(C.c_float), C.c_int32],
    "rcv_sep_filter2d_q8": [_MatP, _MatP, _P(C.c_int32), C.c_int32, _P(C.c_int32), C.c_int32],
    "rcv_filter2d": [_MatP, _MatP, _P(C.c_float), C.c_int32, C.c_int32, C.c_float],
    "rcv_sobel_mag": [_MatP, _MatP, _MatP, _MatP],
    "rcv_resize_bilinear": [_MatP, _MatP],
    "rcv_warp_affine": [_MatP, _MatP, _P(C.c_double), C.c_int32, C.c_double],
    "rcv_get_rotation_matrix_2d": [C.c_double, C.c_double, C.c_double, C.c_double, _P(C.c_double)],
    "rcv_invert_affine": [_P(C.c_double), _P(C.c_double)],
    "rcv_yuyv_to_bgr_gaussian5": [_MatP, _MatP],
    "rcv_yuyv_to_sobel_mag": [_MatP, _MatP],
    "rcv_yuyv_to_sobel_mag_batch": [_MatP, _MatP, C.c_int32],
    "rcv_yuyv_to_bgr_gaussian5_batch": [_MatP, _MatP, C.c_int32],
    "rcv_gaussian_blur_batch": [_MatP, _MatP, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double],
    "rcv_sobel_mag_batch": [_MatP, _MatP, C.c_int32],
    "rcv_resize_bilinear_batch": [_MatP, _MatP, C.c_int32],
    "rcv_warp_affine_batch": [_MatP, _MatP, C.c_int32, _P(C.c_double), C.c_int32, C.c_double],
    "rcv_cvt_color_batch": [_MatP, _MatP, C.c_int32, C.c_int32],
    "rcv_sep_filter2d_q8_batch": [_MatP, _MatP, C.c_int32, _P(C.c_int32), C.c_int32, _P(C.c_int32), C.c_int32],
    "rcv_filter2d_batch": [_MatP, _MatP, C.c_int32, _P(C.c_float), C.c_int32, C.c_int32, C.c_float],
    "rcv_filter2d_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32, _P(C.c_float), C.c_int32, C.c_int32, C.c_float],
    "rcv_gaussian_blur_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double],
    "rcv_sobel_mag_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32],
    "rcv_resize_bilinear_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32],
    "rcv_warp_affine_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32, _P(C.c_double), C.c_int32, C.c_double],
    "rcv_cvt_color_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32, C.c_int32],
    "rcv_yuyv_to_sobel_mag_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32],
    "rcv_yuyv_to_bgr_gaussian5_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32],
    "rcv_sep_filter2d_q8_batch_multi": [_MatP, _MatP, C.c_int32, C.c_int32, _P(C.c_int32), C.c_int32, _P(C.c_int32),
                                        C.c_int32],
    "rcv_set_kernel_broadcast": [_P(C.c_float), C.c_int32, C.c_int32, C.c_int32, _P(C.c_float)],
    "rcv_set_option": [C.c_char_p, C.c_int64],
    "rcv_get_option": [C.c_char_p, _P(C.c_int64)],
}
_RESTYPES = {"rcv_last_error": C.c_char_p, "rcv_version": C.c_char_p}


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m rustcv_b200.build` "
            "(rustcv_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    return lib


lib = load()


def check(rc: int) -> None:
    if rc != RCV_OK:
        raise RcvError(rc, lib.rcv_last_error().decode("utf-8", "replace"))
