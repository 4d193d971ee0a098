"""Device-resident timing of BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py).
One JSON line per config: Mpix/s, algorithmic GB/s, fraction of the measured HBM peak.
GPU box only:  python scripts/bench_configs.py [cfg ...]
"""
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
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
R.imgproc.init(0)
stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(0))
R.imgproc.set_blocking(False)


def fill_batch(batch, make):
    h = None
    for i in range(len(batch)):
        a = make(i)
        h = R.Mat.from_numpy(a)
        F.check(F.lib.rcv_mat_upload(C.byref(h.c()), C.byref(batch[i].c())))


def timeit(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    R.imgproc.sync(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    R.imgproc.sync(0)
    return e0.elapsed_time(e1) / steps


def report(name, ms, units, bytes_per_unit, extra=None):
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    line = {"config": name, "ms_per_step": ms, "units_per_step": units, "Munits_per_s": units / ms / 1e3,
            "algorithmic_bytes_per_unit": bytes_per_unit, "achieved_GBps": gbs, "frac_of_measured_hbm_peak": gbs / PEAK}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def cfg1():
    n, h, w = 256, 480, 640
    src, dst = R.Mat.device_batch(n, h, w, 2), R.Mat.device_batch(n, h, w, 3)
    base = O.fill_u8(1, h * w * 2)
    fill_batch(src, lambda i: np.roll(base, i * 31).reshape(h, w, 2))
    ms = timeit(lambda: R.imgproc.cvt_color_batch(src, dst, R.imgproc.COLOR_YUYV2BGR))
    ok = O.crc32(dst[0].to_numpy()) == 0x0BF66518
    report("cfg1 YUYV->BGR 640x480 x256", ms, n * h * w, 5, {"crc_ok": ok})
    src.free(); dst.free()


def cfg3():
    n, h, w = 64, 1080, 1920
    src, dst = R.Mat.device_batch(n, h, w, 1, R.F32), R.Mat.device_batch(n, h, w, 1, R.F32)
    base = O.fill_f32(3, h * w)
    fill_batch(src, lambda i: np.roll(base, i * 31).reshape(h, w))
    ms = timeit(lambda: R.imgproc.sobel_mag_batch(src, dst))
    O.set_threads(8)
    want = O.sobel3(base.reshape(h, w))["mag"]
    O.set_threads(1)
    ok = bool((np.abs(dst[0].to_numpy() - want) <= 1e-6).all())
    report("cfg3 Sobel3x3+magnitude 1920x1080 f32 x64", ms, n * h * w, 8, {"parity_frame0": ok})
    src.free(); dst.free()


def cfg4():
    n, h, w = 16, 4320, 7680
    src, dst = R.Mat.device_batch(n, h, w, 3), R.Mat.device_batch(n, h // 4, w // 4, 3)
    base = O.fill_u8(4, h * w * 3)
    fill_batch(src, lambda i: (np.roll(base, i * 31) if i else base).reshape(h, w, 3))
    ms = timeit(lambda: R.imgproc.resize_batch(src, dst), steps=10)
    ok = O.crc32(dst[0].to_numpy()) == 0x31A84A85
    report("cfg4 resize 7680x4320->1920x1080 BGR u8 x16", ms, n * (h // 4) * (w // 4), 15,
           {"crc_ok": ok, "sector_floor_bytes_per_unit": 27,
            "frac_vs_sector_floor": n * (h // 4) * (w // 4) * 27 / (ms * 1e-3) / 1e9 / PEAK})
    src.free(); dst.free()


def cfg5():
    n, s = 8, 4096
    src, dst = R.Mat.device_batch(n, s, s, 1, R.F32), R.Mat.device_batch(n, s, s, 1, R.F32)
    base = O.fill_f32(5, s * s)
    fill_batch(src, lambda i: (np.roll(base, i * 31) if i else base).reshape(s, s))
    M = R.imgproc.get_rotation_matrix_2d(((s - 1) / 2, (s - 1) / 2), 15.0)
    ms = timeit(lambda: R.imgproc.warp_affine_batch(src, dst, M), steps=10)
    report("cfg5 warpAffine 15deg 4096x4096 f32 x8", ms, n * s * s, 7.6)
    src.free(); dst.free()


ALL = {"cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}
for name in (sys.argv[1:] or list(ALL)):
    ALL[name]()
