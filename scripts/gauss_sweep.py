"""Device-resident timing of the Gaussian strip kernel under option sweeps (GPU box only).
usage: python scripts/gauss_sweep.py "opt=val,opt=val" "opt=val" ...   (each argument = one configuration)
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rustcv_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from rustcv_b200 import _ffi as F  # noqa: E402

ROWS, COLS, CN, NF = 2160, 3840, 3, int(os.environ.get("NF", "32"))
KS, SIGMA = int(os.environ.get("KS", "5")), float(os.environ.get("SIGMA", "0"))
R.imgproc.init(0)
src = R.Mat.device_batch(NF, ROWS, COLS, CN)
dst = R.Mat.device_batch(NF, ROWS, COLS, CN)
host = R.Mat.pinned(ROWS, COLS, CN)
base = O.fill_u8(2, ROWS * COLS * CN)
for i in range(NF):
    host.data[:] = np.roll(base, i * 7919) if i else base
    F.check(F.lib.rcv_mat_upload(C.byref(host.c()), C.byref(src[i].c())))
stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(0))
R.imgproc.set_blocking(False)
ALL = ["gauss.band_rows", "strip.dynamic", "strip.grid", "gauss.variant", "gauss.no_binomial3"]


def run(cfg: str, steps=20):
    for k in ALL:
        R.imgproc.set_option(k, 1 if k == "strip.dynamic" else 0)
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        R.imgproc.set_option(k, int(v))
    for _ in range(3):
        R.imgproc.gaussian_blur_batch(src, dst, (KS, KS), SIGMA)
    R.imgproc.sync(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        R.imgproc.gaussian_blur_batch(src, dst, (KS, KS), SIGMA)
    e1.record(stream)
    R.imgproc.sync(0)
    ms = e0.elapsed_time(e1) / steps
    ok = (O.crc32(dst[0].to_numpy()) == 0x827081C8) if (KS == 5 and SIGMA == 0) else None
    gbs = 6 * NF * ROWS * COLS / (ms * 1e-3) / 1e9
    print(f"{cfg or 'default':50s} {ms * 1e3 / NF:8.2f} us/frame  {NF * ROWS * COLS / ms / 1e3:10.0f} Mpix/s  "
          f"{gbs:7.0f} GB/s  frac {gbs / 6549.4:.3f}  crc_ok={ok}", flush=True)


for cfg in (sys.argv[1:] or [""]):
    run(cfg)
