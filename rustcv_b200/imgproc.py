"""`rustcv::imgproc` on the B200 backend.

Free functions over `Mat` in the style of the reference's only imgproc functions
(`rectangle(mat, ...)`, rustcv/src/imgproc/drawing.rs:67) with OpenCV argument order:
`op(src, dst, params...)`.  As `VideoCapture::read` does for its output
(rustcv/src/videoio/mod.rs:192-199), the wrapper sizes a HOST `dst` before the call;
device / pinned dsts must already have the right geometry.  Every function forwards to
one entry point of include/rcv_imgproc.h; errors surface as `RcvError` (the Rust
wrapper maps the same codes to `anyhow!`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F
from .mat import F32, U8, Mat, MatBatch

COLOR_YUYV2BGR = F.COLOR_YUYV2BGR
COLOR_UYVY2BGR = F.COLOR_UYVY2BGR
COLOR_BGRA2BGR = F.COLOR_BGRA2BGR
COLOR_RGB2BGR = F.COLOR_RGB2BGR
COLOR_BGR2RGB = F.COLOR_BGR2RGB
COLOR_BGR2GRAY = F.COLOR_BGR2GRAY
COLOR_BGR2XRGB32 = F.COLOR_BGR2XRGB32
COLOR_YUYV2GRAY = F.COLOR_YUYV2GRAY

_DST_CHANNELS = {COLOR_YUYV2BGR: 3, COLOR_UYVY2BGR: 3, COLOR_BGRA2BGR: 3, COLOR_RGB2BGR: 3,
                 COLOR_BGR2GRAY: 1, COLOR_BGR2XRGB32: 4, COLOR_YUYV2GRAY: 1}


def init(device: int = 0) -> None:
    F.check(F.lib.rcv_init(device))


def init_multi(ngpus: int = 0) -> None:
    """GPUs 0..ngpus-1 (0 = every GPU of the box), one library worker thread per GPU."""
    F.check(F.lib.rcv_init_multi(ngpus))


def shutdown() -> None:
    F.check(F.lib.rcv_shutdown())


def _size_dst(dst: Mat, rows: int, cols: int, channels: int, depth: int) -> None:
    if dst.loc == F.RCV_HOST:
        dst.ensure_size(rows, cols, channels, depth)


def cvt_color(src: Mat, dst: Mat, code: int) -> None:
    if code not in _DST_CHANNELS:
        raise F.RcvError(F.RCV_ERR_ARG, f"unknown colour conversion code {code}")
    _size_dst(dst, src.rows, src.cols, _DST_CHANNELS[code], U8)
    F.check(F.lib.rcv_cvt_color(C.byref(src.c()), C.byref(dst.c()), code))


def yuyv_to_bgr(src: Mat, dst: Mat) -> None:
    _size_dst(dst, src.rows, src.cols, 3, U8)
    F.check(F.lib.rcv_yuyv_to_bgr(C.byref(src.c()), C.byref(dst.c())))


def nv12_to_bgr(y: Mat, uv: Mat, dst: Mat) -> None:
    _size_dst(dst, y.rows, y.cols, 3, U8)
    F.check(F.lib.rcv_nv12_to_bgr(C.byref(y.c()), C.byref(uv.c()), C.byref(dst.c())))


def convert_to(src: Mat, dst: Mat, depth: int, alpha: float = 1.0, beta: float = 0.0) -> None:
    """cv::Mat::convertTo between u8 and f32 (dst = saturate(src*alpha + beta))."""
    _size_dst(dst, src.rows, src.cols, src.channels, depth)
    F.check(F.lib.rcv_convert_to(C.byref(src.c()), C.byref(dst.c()), alpha, beta))


def gaussian_blur(src: Mat, dst: Mat, ksize=(5, 5), sigma_x: float = 0.0, sigma_y: float = 0.0) -> None:
    _size_dst(dst, src.rows, src.cols, src.channels, src.depth)
    F.check(F.lib.rcv_gaussian_blur(C.byref(src.c()), C.byref(dst.c()), ksize[0], ksize[1], sigma_x, sigma_y))


def sep_filter2d(src: Mat, dst: Mat, kx, ky) -> None:
    """f32 images take f32 taps; u8 images take Q8 integer taps."""
    _size_dst(dst, src.rows, src.cols, src.channels, src.depth)
    if src.depth == F32:
        kx = np.ascontiguousarray(kx, dtype=np.float32)
        ky = np.ascontiguousarray(ky, dtype=np.float32)
        F.check(F.lib.rcv_sep_filter2d(C.byref(src.c()), C.byref(dst.c()), kx.ctypes.data_as(C.POINTER(C.c_float)),
                                       kx.size, ky.ctypes.data_as(C.POINTER(C.c_float)), ky.size))
    else:
        kx = np.ascontiguousarray(kx, dtype=np.int32)
        ky = np.ascontiguousarray(ky, dtype=np.int32)
        F.check(F.lib.rcv_sep_filter2d_q8(C.byref(src.c()), C.byref(dst.c()), kx.ctypes.data_as(C.POINTER(C.c_int32)),
                                          kx.size, ky.ctypes.data_as(C.POINTER(C.c_int32)), ky.size))


def filter2d(src: Mat, dst: Mat, kernel, delta: float = 0.0) -> None:
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    assert k.ndim == 2
    _size_dst(dst, src.rows, src.cols, src.channels, src.depth)
    F.check(F.lib.rcv_filter2d(C.byref(src.c()), C.byref(dst.c()), k.ctypes.data_as(C.POINTER(C.c_float)),
                               k.shape[1], k.shape[0], delta))


def sobel_mag(src: Mat, mag: Mat | None, gx: Mat | None = None, gy: Mat | None = None) -> None:
    ptrs = []
    for o in (mag, gx, gy):
        if o is None:
            ptrs.append(None)
        else:
            _size_dst(o, src.rows, src.cols, 1, F32)
            ptrs.append(C.byref(o.c()))
    F.check(F.lib.rcv_sobel_mag(C.byref(src.c()), *ptrs))


def resize(src: Mat, dst: Mat, dsize: tuple[int, int] | None = None) -> None:
    """cv::resize INTER_LINEAR; dsize = (width, height) as in OpenCV, or dst's own geometry."""
    if dsize is not None:
        _size_dst(dst, dsize[1], dsize[0], src.channels, src.depth)
    F.check(F.lib.rcv_resize_bilinear(C.byref(src.c()), C.byref(dst.c())))


def get_rotation_matrix_2d(center: tuple[float, float], angle_deg: float, scale: float = 1.0) -> np.ndarray:
    m = (C.c_double * 6)()
    F.check(F.lib.rcv_get_rotation_matrix_2d(center[0], center[1], angle_deg, scale, m))
    return np.array(m[:], dtype=np.float64).reshape(2, 3)


def warp_affine(src: Mat, dst: Mat, M, dsize: tuple[int, int] | None = None, inverse_map: bool = False,
                border_value: float = 0.0) -> None:
    m = (C.c_double * 6)(*[float(v) for v in np.asarray(M, dtype=np.float64).ravel()])
    if dsize is None:
        dsize = (src.cols, src.rows)
    _size_dst(dst, dsize[1], dsize[0], src.channels, src.depth)
    F.check(F.lib.rcv_warp_affine(C.byref(src.c()), C.byref(dst.c()), m, int(inverse_map), border_value))


def yuyv_to_bgr_gaussian5(src: Mat, dst: Mat) -> None:
    _size_dst(dst, src.rows, src.cols, 3, U8)
    F.check(F.lib.rcv_yuyv_to_bgr_gaussian5(C.byref(src.c()), C.byref(dst.c())))


def yuyv_to_sobel_mag(src: Mat, mag: Mat) -> None:
    """sobel_mag(convert_to(cvt_color(cvt_color(src, YUYV2BGR), BGR2GRAY), F32)) in one kernel (bit-identical)."""
    _size_dst(mag, src.rows, src.cols, 1, F32)
    F.check(F.lib.rcv_yuyv_to_sobel_mag(C.byref(src.c()), C.byref(mag.c())))


# ---- batches of independent frames -------------------------------------------------------
def _arr(b) -> tuple:
    if isinstance(b, MatBatch):
        return b.arr, b.n
    mb = MatBatch.of(list(b))
    return mb.arr, mb.n


def gaussian_blur_batch(srcs, dsts, ksize=(5, 5), sigma_x: float = 0.0, sigma_y: float = 0.0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_gaussian_blur_batch(sa, da, n, ksize[0], ksize[1], sigma_x, sigma_y))


def sobel_mag_batch(srcs, mags) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(mags)
    assert n == m
    F.check(F.lib.rcv_sobel_mag_batch(sa, da, n))


def resize_batch(srcs, dsts) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_resize_bilinear_batch(sa, da, n))


def warp_affine_batch(srcs, dsts, M, inverse_map: bool = False, border_value: float = 0.0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    mm = (C.c_double * 6)(*[float(v) for v in np.asarray(M, dtype=np.float64).ravel()])
    F.check(F.lib.rcv_warp_affine_batch(sa, da, n, mm, int(inverse_map), border_value))


def cvt_color_batch(srcs, dsts, code: int) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_cvt_color_batch(sa, da, n, code))


def yuyv_to_sobel_mag_batch(srcs, mags) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(mags)
    assert n == m
    F.check(F.lib.rcv_yuyv_to_sobel_mag_batch(sa, da, n))


def yuyv_to_bgr_gaussian5_batch(srcs, dsts) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_yuyv_to_bgr_gaussian5_batch(sa, da, n))


# ---- the same, sharded over several GPUs from ONE calling thread --------------
def gaussian_blur_batch_multi(srcs, dsts, ngpus: int = 0, ksize=(5, 5), sigma_x: float = 0.0, sigma_y: float = 0.0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_gaussian_blur_batch_multi(sa, da, n, ngpus, ksize[0], ksize[1], sigma_x, sigma_y))


def sobel_mag_batch_multi(srcs, mags, ngpus: int = 0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(mags)
    assert n == m
    F.check(F.lib.rcv_sobel_mag_batch_multi(sa, da, n, ngpus))


def resize_batch_multi(srcs, dsts, ngpus: int = 0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_resize_bilinear_batch_multi(sa, da, n, ngpus))


def warp_affine_batch_multi(srcs, dsts, M, ngpus: int = 0, inverse_map: bool = False, border_value: float = 0.0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    mm = (C.c_double * 6)(*[float(v) for v in np.asarray(M, dtype=np.float64).ravel()])
    F.check(F.lib.rcv_warp_affine_batch_multi(sa, da, n, ngpus, mm, int(inverse_map), border_value))


def cvt_color_batch_multi(srcs, dsts, code: int, ngpus: int = 0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_cvt_color_batch_multi(sa, da, n, ngpus, code))


def yuyv_to_sobel_mag_batch_multi(srcs, mags, ngpus: int = 0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(mags)
    assert n == m
    F.check(F.lib.rcv_yuyv_to_sobel_mag_batch_multi(sa, da, n, ngpus))


def yuyv_to_bgr_gaussian5_batch_multi(srcs, dsts, ngpus: int = 0) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    F.check(F.lib.rcv_yuyv_to_bgr_gaussian5_batch_multi(sa, da, n, ngpus))


def _taps(k):
    if k is None:
        return None, 0
    a = np.ascontiguousarray(k, dtype=np.int32)
    return a, a.size


def sep_filter2d_q8_batch(srcs, dsts, kx, ky) -> None:
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    ax, kw = _taps(kx)
    ay, kh = _taps(ky)
    ip = C.POINTER(C.c_int32)
    F.check(F.lib.rcv_sep_filter2d_q8_batch(sa, da, n, ax.ctypes.data_as(ip), kw, ay.ctypes.data_as(ip), kh))


def filter2d_batch(srcs, dsts, kernel, delta: float = 0.0, ngpus: int = None) -> None:
    """Dense filter2D over a batch (one launch for a uniform device batch); ngpus: fan out over that many GPUs (0 = all)."""
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    kh, kw = k.shape
    kp = k.ctypes.data_as(C.POINTER(C.c_float))
    if ngpus is None:
        F.check(F.lib.rcv_filter2d_batch(sa, da, n, kp, kw, kh, delta))
    else:
        F.check(F.lib.rcv_filter2d_batch_multi(sa, da, n, ngpus, kp, kw, kh, delta))


def sep_filter2d_q8_batch_multi(srcs, dsts, ngpus: int = 0, kx=None, ky=None, kw: int = 0, kh: int = 0) -> None:
    """kx = ky = None: every GPU filters with the kw + kh taps it received from set_kernel_broadcast."""
    sa, n = _arr(srcs)
    da, m = _arr(dsts)
    assert n == m
    ip = C.POINTER(C.c_int32)
    if kx is None:
        F.check(F.lib.rcv_sep_filter2d_q8_batch_multi(sa, da, n, ngpus, None, kw, None, kh))
        return
    ax, kw = _taps(kx)
    ay, kh = _taps(ky)
    F.check(F.lib.rcv_sep_filter2d_q8_batch_multi(sa, da, n, ngpus, ax.ctypes.data_as(ip), kw, ay.ctypes.data_as(ip), kh))


def set_kernel_broadcast(coeffs, root_device: int = 0, ngpus: int = 0) -> np.ndarray:
    """The path's one collective: `coeffs` (f32) from GPU `root_device` to every GPU's coefficient
    bank over NCCL.  Returns what each GPU received, (ngpus, count)."""
    a = np.ascontiguousarray(coeffs, dtype=np.float32).ravel()
    cnt = C.c_int()
    F.check(F.lib.rcv_device_count(C.byref(cnt)))
    got = np.zeros((max(cnt.value, 1), a.size), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    F.check(F.lib.rcv_set_kernel_broadcast(a.ctypes.data_as(fp), a.size, root_device, ngpus, got.ctypes.data_as(fp)))
    return got


# ---- runtime knobs ---------------------------------------------------------------------------
def set_option(name: str, value: int) -> None:
    F.check(F.lib.rcv_set_option(name.encode(), value))


def set_blocking(blocking: bool) -> None:
    F.check(F.lib.rcv_set_blocking(int(blocking)))


def sync(device: int = -1) -> None:
    F.check(F.lib.rcv_sync(device))


def launch_count() -> int:
    n = C.c_uint64()
    F.check(F.lib.rcv_launch_count(C.byref(n)))
    return n.value


def stream_ptr(device: int = -1) -> int:
    p = C.c_void_p()
    F.check(F.lib.rcv_get_stream(device, C.byref(p)))
    return p.value or 0
