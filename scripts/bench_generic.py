"""Device-resident timing of the general-case kernels (not BASELINE configs): GB/s of algorithmic
traffic and fraction of the measured HBM peak.  GPU box only."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rustcv_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from rustcv_b200 import _ffi as F  # noqa: E402

PEAK = 6549.4
R.imgproc.init(0)
stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(0))
R.imgproc.set_blocking(False)
I = R.imgproc
for kv in filter(None, os.environ.get("OPT", "").split(",")):  # OPT=name:value,... sets library options
    I.set_option(kv.split(":")[0], int(kv.split(":")[1]))


def dev(a):
    h = R.Mat.from_numpy(a)
    return h.upload()


def timeit(fn, steps=10, warm=2):
    for _ in range(warm):
        fn()
    I.sync(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    I.sync(0)
    return e0.elapsed_time(e1) / steps


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / PEAK, 3)}), flush=True)


H, W = 2160, 3840
bgr = dev(O.fill_u8(2, H * W * 3).reshape(H, W, 3))
out3 = bgr.like()
want = set(sys.argv[1:])


def case(name):
    return not want or name in want


if case("gauss"):
    for ks, sg in (((3, 3), 0.0), ((7, 7), 1.5), ((5, 5), 1.0), ((9, 9), 1.5), ((11, 11), 0.0), ((13, 13), 2.0), ((15, 15), 0.0),
                   ((31, 31), 5.0)):
        report(f"gauss u8c3 4K {ks} s{sg}", timeit(lambda: I.gaussian_blur(bgr, out3, ks, sg)), 6 * H * W)
    f = dev(O.fill_f32(3, 1080 * 1920).reshape(1080, 1920))
    fo = f.like()
    report("gauss f32c1 1080p 5x5 s1.1", timeit(lambda: I.gaussian_blur(f, fo, (5, 5), 1.1)), 8 * 1080 * 1920)
    report("gauss f32c1 1080p 11x11 s2", timeit(lambda: I.gaussian_blur(f, fo, (11, 11), 2.0)), 8 * 1080 * 1920)
    f3 = dev(O.fill_f32(3, 1080 * 1920 * 3).reshape(1080, 1920, 3))
    fo3 = f3.like()
    report("gauss f32c3 1080p 5x5 s1.1", timeit(lambda: I.gaussian_blur(f3, fo3, (5, 5), 1.1)), 24 * 1080 * 1920)
    report("gauss f32c3 1080p 11x11 s2", timeit(lambda: I.gaussian_blur(f3, fo3, (11, 11), 2.0)), 24 * 1080 * 1920)
if case("gauss11"):
    report("gauss u8c3 4K (11, 11) s0.0", timeit(lambda: I.gaussian_blur(bgr, out3, (11, 11), 0.0)), 6 * H * W)
if case("f2d5"):
    k5 = np.ones((5, 5), np.float32) / 25
    report("filter2d u8c3 4K 5x5", timeit(lambda: I.filter2d(bgr, out3, k5)), 6 * H * W)
if case("filter2d"):
    k = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    report("filter2d u8c3 4K 3x3", timeit(lambda: I.filter2d(bgr, out3, k)), 6 * H * W)
    k5 = np.ones((5, 5), np.float32) / 25
    report("filter2d u8c3 4K 5x5", timeit(lambda: I.filter2d(bgr, out3, k5)), 6 * H * W)
    k7 = np.ones((7, 7), np.float32) / 49
    report("filter2d u8c3 4K 7x7", timeit(lambda: I.filter2d(bgr, out3, k7)), 6 * H * W)
    ff = dev(O.fill_f32(3, 1080 * 1920 * 3).reshape(1080, 1920, 3))
    ffo = ff.like()
    report("filter2d f32c3 1080p 3x3", timeit(lambda: I.filter2d(ff, ffo, k)), 24 * 1080 * 1920)
    report("filter2d f32c3 1080p 5x5", timeit(lambda: I.filter2d(ff, ffo, k5)), 24 * 1080 * 1920)
if case("cvt"):
    g = bgr.like(channels=1)
    report("bgr2gray 4K", timeit(lambda: I.cvt_color(bgr, g, I.COLOR_BGR2GRAY)), 4 * H * W)
    report("rgb2bgr 4K", timeit(lambda: I.cvt_color(bgr, out3, I.COLOR_RGB2BGR)), 6 * H * W)
    x = bgr.like(channels=4)
    report("bgr2xrgb32 4K", timeit(lambda: I.cvt_color(bgr, x, I.COLOR_BGR2XRGB32)), 7 * H * W)
    report("bgra2bgr 4K", timeit(lambda: I.cvt_color(x, out3, I.COLOR_BGRA2BGR)), 7 * H * W)
if case("nv12"):
    yb = R.Mat.device_batch(1, H, W, 1)
    y = dev(O.fill_u8(8, H * W).reshape(H, W))
    uv = dev(O.fill_u8(9, (H // 2) * (W // 2) * 2).reshape(H // 2, W // 2, 2))
    report("nv12->bgr 4K (1 frame/launch)", timeit(lambda: I.nv12_to_bgr(y, uv, out3), steps=50), int(4.5 * H * W))
    big = dev(O.fill_u8(8, 4 * H * W).reshape(4 * H, W))
    buv = dev(O.fill_u8(9, 4 * (H // 2) * (W // 2) * 2).reshape(2 * H, W // 2, 2))
    bout = big.like(channels=3)
    report("nv12->bgr 3840x8640 (4 stacked 4K frames)", timeit(lambda: I.nv12_to_bgr(big, buv, bout), steps=20), int(4.5 * 4 * H * W))
if case("resize"):
    # bytes = the source rows a bilinear kernel must touch (whole rows: every 32-byte sector of a used row is hit
    # for scales below ~8) + the destination; "tile" = shared-memory staged kernel, "naive" = per-tap global loads
    for name, dr, dc, rows_used in (("4K->720p (3x)", 720, 1280, 2 * 720), ("4K->1080p (2x)", 1080, 1920, H),
                                    ("4K->8K (upscale 2x)", 4320, 7680, H), ("4K->1600x900 (2.4x)", 900, 1600, 2 * 900)):
        d = bgr.like(rows=dr, cols=dc)
        nbytes = rows_used * W * 3 + dr * dc * 3
        for opt, label in ((0, "tile"), (1, "naive")):
            I.set_option("resize.force_generic", opt)
            report(f"resize u8c3 {name} {label}", timeit(lambda: I.resize(bgr, d)), nbytes)
        I.set_option("resize.force_generic", 0)
    f4 = dev(O.fill_f32(4, 2160 * 3840).reshape(2160, 3840))
    fd = f4.like(rows=1080, cols=1920)
    for opt, label in ((0, "tile"), (1, "naive")):
        I.set_option("resize.force_generic", opt)
        report(f"resize f32c1 4K->1080p {label}", timeit(lambda: I.resize(f4, fd)), 4 * (H * W + 1080 * 1920))
    I.set_option("resize.force_generic", 0)
if case("warp"):
    M = I.get_rotation_matrix_2d(((W - 1) / 2, (H - 1) / 2), 15.0)
    report("warpAffine u8c3 4K 15deg", timeit(lambda: I.warp_affine(bgr, out3, M)), 6 * H * W)
if case("single"):
    report("gauss5 single 4K frame (1 launch)", timeit(lambda: I.gaussian_blur(bgr, out3, (5, 5), 0.0), steps=50), 6 * H * W)
