"""ctypes loader for the CPU oracle (oracle/librcv_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/rcv_oracle.h.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the rustcv_b200 package.

Arrays are numpy; any array whose last axes are C-contiguous within a row is
accepted, the row pitch (``step`` in rustcv::core::Mat, rustcv/src/core/mat.rs:12)
is taken from ``arr.strides[0]``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librcv_oracle.so")

GAMMA = np.uint64(0x9E3779B97F4A7C15)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rcv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_crc32.restype = C.c_uint32
        _lib.orc_gaussian_ksize.restype = C.c_int
        _lib.orc_get_threads.restype = C.c_int
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(int(n)))


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _step(a: np.ndarray) -> C.c_size_t:
    return C.c_size_t(a.strides[0] if a.ndim >= 2 else a.nbytes)


def _check_rows(a: np.ndarray):
    assert a.ndim in (2, 3)
    inner = a[0]
    assert inner.flags["C_CONTIGUOUS"] or inner.size <= 1, "rows must be contiguous"


# --------------------------------------------------------------------------
# synthetic data (vectorised numpy SplitMix64; bit-identical to orc_fill_*)
# --------------------------------------------------------------------------
def splitmix64(seed: int, n: int) -> np.ndarray:
    """First ``n`` outputs of SplitMix64(seed) as uint64."""
    with np.errstate(over="ignore"):
        s = np.uint64(seed) + GAMMA * np.arange(1, n + 1, dtype=np.uint64)
        z = s
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def fill_u8(seed: int, n: int) -> np.ndarray:
    z = splitmix64(seed, (n + 7) // 8)
    return z.astype("<u8").view(np.uint8)[:n].copy()


def fill_f32(seed: int, n: int) -> np.ndarray:
    z = splitmix64(seed, (n + 1) // 2)
    v = z.astype("<u8").view("<u4")[:n]
    return ((v >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def crc32(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def padded(a: np.ndarray, step: int, fill: int = 0xA5) -> np.ndarray:
    """Copy ``a`` (rows x ...) into a buffer whose row pitch is ``step`` bytes and
    return a strided view of the valid part (exercises Mat.step > cols*channels)."""
    rows = a.shape[0]
    row_bytes = a[0].nbytes
    assert step >= row_bytes and step % a.itemsize == 0
    buf = np.full((rows, step), fill, dtype=np.uint8)
    buf[:, :row_bytes] = np.ascontiguousarray(a).view(np.uint8).reshape(rows, row_bytes)
    view = buf[:, :row_bytes].view(a.dtype).reshape((rows,) + a.shape[1:])
    return view


# --------------------------------------------------------------------------
# conversions
# --------------------------------------------------------------------------
def yuyv_to_bgr_facade(src: np.ndarray, width: int, height: int, dst_len: int | None = None):
    """rustcv/src/videoio/mod.rs:344-371 on flat byte buffers.
    Returns (status, dst) -- status 0 converted, 1 silent return, -1 would panic."""
    src = np.ascontiguousarray(src, dtype=np.uint8).ravel()
    n = width * height * 3 if dst_len is None else dst_len
    dst = np.zeros(n, dtype=np.uint8)
    lib().orc_yuyv_to_bgr_facade.restype = C.c_int
    st = lib().orc_yuyv_to_bgr_facade(_p(src), C.c_size_t(src.size), _p(dst), C.c_size_t(dst.size),
                                      C.c_size_t(width), C.c_size_t(height))
    return st, dst


def yuyv_to_bgr_camera(src: np.ndarray, width: int, height: int, dst_len: int | None = None):
    """rustcv-camera/src/decode.rs:160-191."""
    src = np.ascontiguousarray(src, dtype=np.uint8).ravel()
    n = width * height * 3 if dst_len is None else dst_len
    dst = np.zeros(n, dtype=np.uint8)
    lib().orc_yuyv_to_bgr_camera.restype = C.c_int
    st = lib().orc_yuyv_to_bgr_camera(_p(src), C.c_size_t(src.size), _p(dst), C.c_size_t(dst.size),
                                      C.c_size_t(width), C.c_size_t(height))
    return st, dst


def bgra_to_bgr_facade(src: np.ndarray, width: int, height: int, dst_len: int | None = None):
    src = np.ascontiguousarray(src, dtype=np.uint8).ravel()
    n = width * height * 3 if dst_len is None else dst_len
    dst = np.zeros(n, dtype=np.uint8)
    lib().orc_bgra_to_bgr_facade.restype = C.c_int
    st = lib().orc_bgra_to_bgr_facade(_p(src), C.c_size_t(src.size), _p(dst), C.c_size_t(dst.size),
                                      C.c_size_t(width), C.c_size_t(height))
    return st, dst


def _cvt(fn_name, src, out_shape, out_dtype=np.uint8, cols=None):
    _check_rows(src)
    rows = src.shape[0]
    dst = np.zeros(out_shape, dtype=out_dtype)
    getattr(lib(), fn_name)(_p(src), _step(src), _p(dst), _step(dst), C.c_int(rows), C.c_int(cols))
    return dst


def yuyv_to_bgr(src: np.ndarray) -> np.ndarray:
    """src: rows x cols x 2 u8 (YUYV) -> rows x cols x 3 BGR."""
    rows, cols = src.shape[:2]
    return _cvt("orc_yuyv_to_bgr_strided", src, (rows, cols, 3), cols=cols)


def uyvy_to_bgr(src: np.ndarray) -> np.ndarray:
    rows, cols = src.shape[:2]
    return _cvt("orc_uyvy_to_bgr_strided", src, (rows, cols, 3), cols=cols)


def yuyv_to_gray(src: np.ndarray) -> np.ndarray:
    rows, cols = src.shape[:2]
    return _cvt("orc_yuyv_to_gray_strided", src, (rows, cols), cols=cols)


def nv12_to_bgr(y: np.ndarray, uv: np.ndarray) -> np.ndarray:
    rows, cols = y.shape[:2]
    dst = np.zeros((rows, cols, 3), dtype=np.uint8)
    lib().orc_nv12_to_bgr_strided(_p(y), _step(y), _p(uv), _step(uv), _p(dst), _step(dst),
                                  C.c_int(rows), C.c_int(cols))
    return dst


def bgra_to_bgr(src: np.ndarray) -> np.ndarray:
    rows, cols = src.shape[:2]
    return _cvt("orc_bgra_to_bgr_strided", src, (rows, cols, 3), cols=cols)


def swap_rb(src: np.ndarray) -> np.ndarray:
    rows, cols = src.shape[:2]
    return _cvt("orc_swap_rb_strided", src, (rows, cols, 3), cols=cols)


def bgr_to_gray(src: np.ndarray) -> np.ndarray:
    rows, cols = src.shape[:2]
    return _cvt("orc_bgr_to_gray_strided", src, (rows, cols), cols=cols)


def convert_to(src: np.ndarray, dtype, alpha: float = 1.0, beta: float = 0.0) -> np.ndarray:
    """cv::Mat::convertTo between u8 and f32."""
    _check_rows(src)
    dst = np.zeros(src.shape, dtype=dtype)
    ncols = src.shape[1] * _cn(src)
    lib().orc_convert_to(_p(src), _step(src), C.c_int(int(src.dtype == np.float32)), _p(dst), _step(dst),
                         C.c_int(int(np.dtype(dtype) == np.float32)), C.c_int(src.shape[0]), C.c_int(ncols),
                         C.c_double(alpha), C.c_double(beta))
    return dst


def bgr_to_xrgb32(src: np.ndarray) -> np.ndarray:
    rows, cols = src.shape[:2]
    return _cvt("orc_bgr_to_xrgb32_strided", src, (rows, cols), out_dtype=np.uint32, cols=cols)


# --------------------------------------------------------------------------
# filters
# --------------------------------------------------------------------------
def gaussian_kernel_q8(n: int, sigma: float) -> np.ndarray:
    k = (C.c_int * 64)()
    lib().orc_gaussian_kernel_q8(C.c_int(n), C.c_double(sigma), k)
    return np.array(k[:n], dtype=np.int32)


def gaussian_kernel_f64(n: int, sigma: float) -> np.ndarray:
    k = (C.c_double * 64)()
    lib().orc_gaussian_kernel_f64(C.c_int(n), C.c_double(sigma), k)
    return np.array(k[:n], dtype=np.float64)


def gaussian_ksize(sigma: float, is_u8: bool) -> int:
    return lib().orc_gaussian_ksize(C.c_double(sigma), C.c_int(int(is_u8)))


def _cn(a: np.ndarray) -> int:
    return 1 if a.ndim == 2 else a.shape[2]


def sepfilter_u8_q8(src, kx, ky) -> np.ndarray:
    _check_rows(src)
    kx = np.ascontiguousarray(kx, dtype=np.int32)
    ky = np.ascontiguousarray(ky, dtype=np.int32)
    dst = np.zeros(src.shape, dtype=np.uint8)
    lib().orc_sepfilter_u8_q8(_p(src), _step(src), _p(dst), _step(dst), C.c_int(src.shape[0]),
                              C.c_int(src.shape[1]), C.c_int(_cn(src)), _p(kx), C.c_int(kx.size),
                              _p(ky), C.c_int(ky.size))
    return dst


def gaussian_blur(src, ksize=(5, 5), sigma_x=0.0, sigma_y=0.0) -> np.ndarray:
    _check_rows(src)
    dst = np.zeros(src.shape, dtype=src.dtype)
    fn = {np.dtype(np.uint8): "orc_gaussian_blur_u8", np.dtype(np.float32): "orc_gaussian_blur_f32"}[src.dtype]
    getattr(lib(), fn)(_p(src), _step(src), _p(dst), _step(dst), C.c_int(src.shape[0]),
                       C.c_int(src.shape[1]), C.c_int(_cn(src)), C.c_int(ksize[0]), C.c_int(ksize[1]),
                       C.c_double(sigma_x), C.c_double(sigma_y))
    return dst


def gaussian5_fast(src) -> np.ndarray:
    """Tuned restatement of gaussian_blur(src, (5, 5), 0): bit-identical, auto-vectorised (bench.py's second CPU figure)."""
    _check_rows(src)
    dst = np.zeros_like(src)
    lib().orc_gauss5_binomial_u8_fast(_p(src), _step(src), _p(dst), _step(dst), C.c_int(src.shape[0]),
                                      C.c_int(src.shape[1]), C.c_int(_cn(src)))
    return dst


def sepfilter_f32(src, kx, ky) -> np.ndarray:
    _check_rows(src)
    kx = np.ascontiguousarray(kx, dtype=np.float32)
    ky = np.ascontiguousarray(ky, dtype=np.float32)
    dst = np.zeros(src.shape, dtype=np.float32)
    lib().orc_sepfilter_f32(_p(src), _step(src), _p(dst), _step(dst), C.c_int(src.shape[0]),
                            C.c_int(src.shape[1]), C.c_int(_cn(src)), _p(kx), C.c_int(kx.size),
                            _p(ky), C.c_int(ky.size))
    return dst


def filter2d(src, kernel, delta=0.0) -> np.ndarray:
    _check_rows(src)
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    dst = np.zeros(src.shape, dtype=src.dtype)
    fn = {np.dtype(np.uint8): "orc_filter2d_u8", np.dtype(np.float32): "orc_filter2d_f32"}[src.dtype]
    getattr(lib(), fn)(_p(src), _step(src), _p(dst), _step(dst), C.c_int(src.shape[0]),
                       C.c_int(src.shape[1]), C.c_int(_cn(src)), _p(k), C.c_int(k.shape[1]),
                       C.c_int(k.shape[0]), C.c_float(delta))
    return dst


def sobel3(src, want=("mag",)):
    """Returns dict with any of gx, gy, mag (f32, single channel)."""
    _check_rows(src)
    assert src.ndim == 2 and src.dtype == np.float32
    out = {k: np.zeros(src.shape, dtype=np.float32) for k in want}

    def arg(k):
        if k in out:
            return _p(out[k]), _step(out[k])
        return C.c_void_p(0), C.c_size_t(0)

    gx, gxs = arg("gx")
    gy, gys = arg("gy")
    mg, mgs = arg("mag")
    lib().orc_sobel3_f32(_p(src), _step(src), gx, gxs, gy, gys, mg, mgs, C.c_int(src.shape[0]),
                         C.c_int(src.shape[1]))
    return out


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def resize_bilinear(src, drows: int, dcols: int) -> np.ndarray:
    _check_rows(src)
    cn = _cn(src)
    shape = (drows, dcols) if src.ndim == 2 else (drows, dcols, cn)
    dst = np.zeros(shape, dtype=src.dtype)
    fn = {np.dtype(np.uint8): "orc_resize_bilinear_u8", np.dtype(np.float32): "orc_resize_bilinear_f32"}[src.dtype]
    getattr(lib(), fn)(_p(src), _step(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), _p(dst),
                       _step(dst), C.c_int(drows), C.c_int(dcols), C.c_int(cn))
    return dst


def rotation_matrix(cx, cy, angle_deg, scale=1.0) -> np.ndarray:
    m = (C.c_double * 6)()
    lib().orc_rotation_matrix(C.c_double(cx), C.c_double(cy), C.c_double(angle_deg), C.c_double(scale), m)
    return np.array(m[:], dtype=np.float64)


def invert_affine(M) -> np.ndarray:
    m = (C.c_double * 6)(*[float(v) for v in np.asarray(M).ravel()])
    im = (C.c_double * 6)()
    lib().orc_invert_affine.restype = C.c_int
    assert lib().orc_invert_affine(m, im) == 0
    return np.array(im[:], dtype=np.float64)


def warp_affine(src, M, dsize=None, inverse_map=False, border_value=0) -> np.ndarray:
    _check_rows(src)
    drows, dcols = (src.shape[0], src.shape[1]) if dsize is None else dsize
    m = (C.c_double * 6)(*[float(v) for v in np.asarray(M).ravel()])
    cn = _cn(src)
    shape = (drows, dcols) if src.ndim == 2 else (drows, dcols, cn)
    dst = np.zeros(shape, dtype=src.dtype)
    if src.dtype == np.float32:
        assert cn == 1
        lib().orc_warp_affine_f32(_p(src), _step(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]),
                                  _p(dst), _step(dst), C.c_int(drows), C.c_int(dcols), m,
                                  C.c_int(int(inverse_map)), C.c_float(border_value))
    else:
        lib().orc_warp_affine_u8(_p(src), _step(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]),
                                 _p(dst), _step(dst), C.c_int(drows), C.c_int(dcols), C.c_int(cn), m,
                                 C.c_int(int(inverse_map)), C.c_int(int(border_value)))
    return dst
