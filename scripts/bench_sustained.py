"""The metric step (32 x 4K BGR GaussianBlur 5x5, one launch) looped for ~1.5 s per setting under the clock sampler:
what the kernel sustains once the GPU sits at its power cap.  GPU box only.
    python scripts/bench_sustained.py name=value[,name=value...] ...      (each argument = one setting of library options)
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rustcv_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

SECONDS = float(os.environ.get("SECONDS_PER_SETTING", "1.5"))
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.0
R.imgproc.init(0)
stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(0))
R.imgproc.set_blocking(False)
F_, ROWS, COLS, CN = 32, 2160, 3840, 3
src, dst = R.Mat.device_batch(F_, ROWS, COLS, CN), R.Mat.device_batch(F_, ROWS, COLS, CN)
base = O.fill_u8(2, ROWS * COLS * CN)
import ctypes as C
from rustcv_b200 import _ffi as F
h = R.Mat.from_numpy(base.reshape(ROWS, COLS, CN))
for i in range(F_):
    F.check(F.lib.rcv_mat_upload(C.byref(h.c()), C.byref(src[i].c())))
sampler = bench.ClockSampler(0)
sampler.start()
OP = os.environ.get("OP", "gauss5")
NF = int(os.environ.get("NF", "64"))
GB = 6 * F_ * ROWS * COLS / 1e9  # algorithmic bytes per step
if OP == "sobel":
    src.free(); dst.free()
    src, dst = R.Mat.device_batch(NF, 1080, 1920, 1, R.F32), R.Mat.device_batch(NF, 1080, 1920, 1, R.F32)
    hf = R.Mat.from_numpy(O.fill_f32(3, 1080 * 1920).reshape(1080, 1920))
    for i in range(NF):
        F.check(F.lib.rcv_mat_upload(C.byref(hf.c()), C.byref(src[i].c())))
    GB = 8 * NF * 1080 * 1920 / 1e9
if OP == "warp":
    src.free(); dst.free()
    NF = int(os.environ.get("NF", "16"))
    src, dst = R.Mat.device_batch(NF, 4096, 4096, 1, R.F32), R.Mat.device_batch(NF, 4096, 4096, 1, R.F32)
    hf = R.Mat.from_numpy(O.fill_f32(5, 4096 * 4096).reshape(4096, 4096))
    for i in range(NF):
        F.check(F.lib.rcv_mat_upload(C.byref(hf.c()), C.byref(src[i].c())))
    GB = 7.6 * NF * 4096 * 4096 / 1e9
    WM = R.imgproc.get_rotation_matrix_2d(((4096 - 1) / 2, (4096 - 1) / 2), 15.0)


def step():
    if OP == "gauss5":
        R.imgproc.gaussian_blur_batch(src, dst, (5, 5), 0.0, 0.0)
    elif OP == "gauss3":
        R.imgproc.gaussian_blur_batch(src, dst, (3, 3), 0.0, 0.0)
    elif OP == "gaussq5":
        R.imgproc.gaussian_blur_batch(src, dst, (5, 5), 1.0, 1.0)
    elif OP == "swap":
        R.imgproc.cvt_color_batch(src, dst, R.imgproc.COLOR_RGB2BGR)
    elif OP == "sobel":
        R.imgproc.sobel_mag_batch(src, dst)
    elif OP == "warp":
        R.imgproc.warp_affine_batch(src, dst, WM)


for setting in (sys.argv[1:] or ["default"]):
    opts = {}
    if setting != "default":
        for kv in setting.split(","):
            k, v = kv.split("=")
            opts[k] = int(v)
    for k, v in opts.items():
        R.imgproc.set_option(k, v)
    for _ in range(5):
        step()
    R.imgproc.sync(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # burst: 20 steps after a pause
    time.sleep(1.0)
    e0.record(stream)
    for _ in range(20):
        step()
    e1.record(stream)
    R.imgproc.sync(0)
    burst_ms = e0.elapsed_time(e1) / 20
    n = int(math.ceil(SECONDS * 1e3 / burst_ms))
    t0 = time.time()
    e0.record(stream)
    for _ in range(n):
        step()
    e1.record(stream)
    R.imgproc.sync(0)
    t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    clk = sampler.window(t0 + 0.3, t1)
    ok = O.crc32(dst[0].to_numpy()) == 0x827081C8 if OP == "gauss5" else None
    gb = GB
    print(json.dumps({"op": OP, "setting": setting, "burst_ms": round(burst_ms, 4), "burst_frac": round(gb / (burst_ms * 1e-3) / PEAK, 4),
                      "sustained_ms": round(ms, 4), "sustained_frac": round(gb / (ms * 1e-3) / PEAK, 4), "sm_mhz": clk.get("sm_mhz"),
                      "power_w_max": clk.get("power_w_max"), "reasons": clk.get("reasons"), "parity": ok}), flush=True)
    for k in opts:
        R.imgproc.set_option(k, 0)
    time.sleep(1.0)
sampler.stop()
