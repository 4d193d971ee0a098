"""Does the content of the pixels move the power-capped (sustained) figures?  HBM / L2 / register-file power depends on
how many data lines toggle: white noise is the worst case, constant data the best (MEASURED_PEAKS.json's copy ran on
whatever torch.empty held).  Runs a plain torch copy and the metric step (32 x 4K BGR GaussianBlur 5x5) for ~2 s each on
three contents -- SplitMix64 noise (the bench input), a photo-like synthetic frame (smooth gradients + +-3 levels of
noise), zeros -- under the clock sampler.  GPU box only.
    python scripts/bench_power_data.py
"""
import ctypes as C
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
from rustcv_b200 import _ffi as F  # noqa: E402

SECONDS = float(os.environ.get("SECONDS_PER_SETTING", "2.0"))
PEAK = 6551.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])
R.imgproc.init(0)
stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(0))
R.imgproc.set_blocking(False)
F_, ROWS, COLS, CN = 32, 2160, 3840, 3
src, dst = R.Mat.device_batch(F_, ROWS, COLS, CN), R.Mat.device_batch(F_, ROWS, COLS, CN)


def content(kind, seed):
    if kind == "noise":
        return O.fill_u8(seed, ROWS * COLS * CN).reshape(ROWS, COLS, CN)
    if kind == "zeros":
        return np.zeros((ROWS, COLS, CN), np.uint8)
    # photo-like: low-frequency gradients per channel + sensor noise of a few levels
    y, x = np.mgrid[0:ROWS, 0:COLS].astype(np.float32)
    rng = np.random.default_rng(seed)
    img = np.empty((ROWS, COLS, CN), np.float32)
    for c in range(CN):
        img[..., c] = 128 + 90 * np.sin(x / (300 + 70 * c) + seed) * np.cos(y / (220 + 50 * c)) + rng.normal(0, 1.5, (ROWS, COLS))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


sampler = bench.ClockSampler(0)
sampler.start()
GB = 6 * F_ * ROWS * COLS / 1e9
for kind in ("noise", "photo", "zeros"):
    for i in range(F_):
        h = R.Mat.from_numpy(content(kind, 2 + (i % 4)))
        F.check(F.lib.rcv_mat_upload(C.byref(h.c()), C.byref(src[i].c())))
    R.imgproc.sync(0)
    # torch views of the same device memory for the plain copy
    nbytes = F_ * ROWS * COLS * CN
    a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    hostall = np.concatenate([content(kind, 2 + (i % 4)).reshape(-1) for i in range(4)])
    a.copy_(torch.from_numpy(np.tile(hostall, F_ // 4)))
    torch.cuda.synchronize()
    for what in ("torch_copy", "gauss5"):
        if what == "gauss5":
            def step():
                R.imgproc.gaussian_blur_batch(src, dst, (5, 5), 0.0, 0.0)
            st = stream
        else:
            def step():
                b.copy_(a)
            st = torch.cuda.current_stream()
        for _ in range(5):
            step()
        torch.cuda.synchronize(); R.imgproc.sync(0)
        time.sleep(1.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            step()
        e1.record(st)
        torch.cuda.synchronize(); R.imgproc.sync(0)
        burst_ms = e0.elapsed_time(e1) / 20
        n = int(math.ceil(SECONDS * 1e3 / burst_ms))
        t0 = time.time()
        e0.record(st)
        for _ in range(n):
            step()
        e1.record(st)
        torch.cuda.synchronize(); R.imgproc.sync(0)
        t1 = time.time()
        ms = e0.elapsed_time(e1) / n
        clk = sampler.window(t0 + 0.3, t1)
        print(json.dumps({"content": kind, "what": what, "burst_ms": round(burst_ms, 4), "burst_gbs": round(GB / (burst_ms * 1e-3)),
                          "burst_frac": round(GB / (burst_ms * 1e-3) / PEAK, 4), "sustained_ms": round(ms, 4),
                          "sustained_gbs": round(GB / (ms * 1e-3)), "sustained_frac": round(GB / (ms * 1e-3) / PEAK, 4),
                          "sm_mhz": clk.get("sm_mhz"), "power_w_max": clk.get("power_w_max"), "reasons": clk.get("reasons")}), flush=True)
        time.sleep(1.5)
    del a, b
sampler.stop()
