"""The pixel-format dispatch of `VideoCapture::read` / `decode_frame`, on the GPU.

Mirrors rustcv/src/videoio/mod.rs:181-260 (size the Mat, branch on FourCC, convert) and
rustcv-camera/src/decode.rs:36-86.  Camera I/O itself (drivers, streams) is out
of scope; this module is the step between a raw (or MJPG) frame buffer and
a BGR `Mat`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F
from .mat import U8, Mat


def fourcc(a: str) -> int:
    """rustcv-core/src/pixel_format.rs:6-33: little-endian packed four characters."""
    b = a.encode("ascii")
    assert len(b) == 4
    return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24)


YUYV = fourcc("YUYV")
UYVY = fourcc("UYVY")
BGRA = fourcc("BGRA")
RGB3 = fourcc("RGB3")
BGR3 = fourcc("BGR3")
NV12 = fourcc("NV12")
MJPEG = fourcc("MJPG")


def mjpeg_info(data: np.ndarray) -> tuple[int, int]:
    """(width, height) from the JPEG header (Decompressor::read_header, videoio/mod.rs:214-216)."""
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    w, h = C.c_int32(0), C.c_int32(0)
    F.check(F.lib.rcv_mjpeg_info(data.ctypes.data, data.size, C.byref(w), C.byref(h)))
    return w.value, h.value


def decode_frame(data: np.ndarray, width: int, height: int, fcc: int, mat: Mat, stride: int | None = None) -> bool:
    """Converts one raw frame into `mat` (BGR, 3 channels), sizing it like `read` does.

    Packed input (`stride` None) follows the facade: stride ignored, width*height/2
    macro-pixels (videoio/mod.rs:203,344-371).  With `stride` the rows are `stride`
    bytes apart (Frame.stride, rustcv-core/src/frame.rs:20-22 -- populated by the
    backends and ignored by the reference's facade).
    MJPG frames (videoio/mod.rs:205-232) are decoded by nvJPEG straight to BGR at the Mat's pitch; width and
    height come from the JPEG header, as in the reference.  Returns False for formats this path does not convert.
    """
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    if fcc == MJPEG:
        w, h = mjpeg_info(data)
        if mat.loc == F.RCV_HOST:
            mat.ensure_size(h, w, 3, U8)
        F.check(F.lib.rcv_mjpeg_to_bgr(data.ctypes.data, data.size, C.byref(mat.c())))
        return True
    if mat.loc == F.RCV_HOST:
        mat.ensure_size(height, width, 3, U8)  # videoio/mod.rs:192-199
    elif (mat.rows, mat.cols, mat.channels) != (height, width, 3):
        raise F.RcvError(F.RCV_ERR_SIZE, "a device / pinned Mat must already be height x width x 3")
    # A device-resident `mat`: only the RAW frame crosses PCIe (2 B/px for
    # YUYV instead of 3 B/px of BGR), the conversion runs on the GPU and the BGR stays in HBM for the
    # imgproc calls that follow.  Row-wise form; identical to the packed run whenever width is even.
    packed_ok = mat.loc != F.RCV_DEVICE and mat.step == width * 3
    if fcc in (YUYV, BGRA) and stride is None and packed_ok:
        fn = F.lib.rcv_yuyv_to_bgr_packed if fcc == YUYV else F.lib.rcv_bgra_to_bgr_packed
        F.check(fn(data.ctypes.data, data.size, mat.data.ctypes.data, mat.data.size, width, height))
        return True
    if fcc == YUYV and stride is None and (width & 1):
        raise F.RcvError(F.RCV_ERR_UNSUPPORTED, "odd-width packed YUYV pairs straddle rows: use a packed host Mat")
    bpp = {YUYV: 2, UYVY: 2, BGRA: 4, RGB3: 3, BGR3: 3}.get(fcc)
    if bpp is None:
        return False
    step = stride if stride is not None else width * bpp
    if data.size < (height - 1) * step + width * bpp:
        raise F.RcvError(F.RCV_ERR_SIZE, "frame buffer shorter than height x stride")
    src = Mat()
    src.data, src.rows, src.cols, src.step, src.channels, src.depth = data, height, width, step, bpp, U8
    if fcc == BGR3:  # "Assume RGB/BGR or copy" (videoio/mod.rs:253-257)
        if mat.loc == F.RCV_DEVICE:
            F.check(F.lib.rcv_mat_upload(C.byref(src.c()), C.byref(mat.c())))
            return True
        for r in range(height):
            mat.row_bytes(r)[:] = src.row_bytes(r)
        return True
    code = {YUYV: F.COLOR_YUYV2BGR, UYVY: F.COLOR_UYVY2BGR, BGRA: F.COLOR_BGRA2BGR, RGB3: F.COLOR_RGB2BGR}[fcc]
    F.check(F.lib.rcv_cvt_color(C.byref(src.c()), C.byref(mat.c()), code))
    return True
