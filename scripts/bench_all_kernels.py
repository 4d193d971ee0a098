"""Every kernel family on batches of device-resident frames: burst (20 steps after a pause) and sustained (>= 0.5 s)
fraction of the measured HBM roofline, each with the nvidia-smi clock / throttle sample of its own sustained run.
GPU box only.  One JSON line per case."""
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

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.0
SECONDS = float(os.environ.get("SECONDS_PER_CASE", "0.5"))
I = R.imgproc
I.init(0)
stream = torch.cuda.ExternalStream(I.stream_ptr(0))
I.set_blocking(False)
sampler = bench.ClockSampler(0)
sampler.start()
want = set(sys.argv[1:])


def batch(n, h, w, cn, depth=R.U8, seed=1):
    b = R.Mat.device_batch(n, h, w, cn, depth)
    if depth == R.U8:
        a = O.fill_u8(seed, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    else:
        a = O.fill_f32(seed, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    hm = R.Mat.from_numpy(a)
    for i in range(n):
        F.check(F.lib.rcv_mat_upload(C.byref(hm.c()), C.byref(b[i].c())))
    return b


def run(name, fn, nbytes, note=""):
    if want and not any(w in name for w in want):
        return
    for _ in range(3):
        fn()
    I.sync(0)
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        fn()
    e1.record(stream)
    I.sync(0)
    burst = e0.elapsed_time(e1) / 20
    n = max(20, int(math.ceil(SECONDS * 1e3 / burst)))
    t0 = time.time()
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    I.sync(0)
    t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    clk = sampler.window(t0 + 0.1, t1)
    print(json.dumps({"case": name, "burst_ms": round(burst, 4), "burst_frac": round(nbytes / (burst * 1e-3) / 1e9 / PEAK, 3),
                      "sustained_ms": round(ms, 4), "sustained_frac": round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 3),
                      "sm_mhz": clk.get("sm_mhz"), "power_w_max": clk.get("power_w_max"), "reasons": clk.get("reasons"), "note": note}), flush=True)


H, W, N = 2160, 3840, 32
PX = N * H * W
bgr = batch(N, H, W, 3)
out3 = R.Mat.device_batch(N, H, W, 3)
run("GaussianBlur 5x5 sigma 0, 4K BGR u8 x32 [k_strip<Gauss5Op<3>>]", lambda: I.gaussian_blur_batch(bgr, out3, (5, 5), 0.0, 0.0), 6 * PX)
run("GaussianBlur 3x3 sigma 0, 4K BGR u8 x32 [k_strip<Gauss3Op<3>>]", lambda: I.gaussian_blur_batch(bgr, out3, (3, 3), 0.0, 0.0), 6 * PX)
for ks, sg in ((3, 0.8), (5, 1.0), (7, 1.5)):
    run(f"GaussianBlur {ks}x{ks} sigma {sg}, 4K BGR u8 x32 [k_strip<GaussQ8Op<3,{ks}>>]", lambda: I.gaussian_blur_batch(bgr, out3, (ks, ks), sg, sg), 6 * PX)
for ks, sg in ((9, 1.5), (11, 2.0), (13, 2.0), (15, 2.5)):
    opn = "GaussQ8Op" if ks <= 11 else "GaussQ8WideOp"
    run(f"GaussianBlur {ks}x{ks} sigma {sg}, 4K BGR u8 x32 [k_strip<{opn}<3,{ks}>>]", lambda: I.gaussian_blur_batch(bgr, out3, (ks, ks), sg, sg), 6 * PX)
    if ks <= 11:
        I.set_option("gauss.wide_windowed", 1)
        run(f"GaussianBlur {ks}x{ks} sigma {sg}, 4K BGR u8 x32, windowed wide op [k_strip<GaussQ8WideOp<3,{ks}>>]", lambda: I.gaussian_blur_batch(bgr, out3, (ks, ks), sg, sg), 6 * PX)
        I.set_option("gauss.wide_windowed", 0)
I.set_option("gauss.no_wide", 1)
run("GaussianBlur 11x11 sigma 2, 4K BGR u8 x32, general kernel [k_sepfilter<u8,11>]", lambda: I.gaussian_blur_batch(bgr, out3, (11, 11), 2.0, 2.0), 6 * PX)
I.set_option("gauss.no_wide", 0)
run("GaussianBlur 17x17 sigma 3, 4K BGR u8 x32, general kernel [k_sepfilter<u8,0>]", lambda: I.gaussian_blur_batch(bgr, out3, (17, 17), 3.0, 3.0), 6 * PX)
g1 = R.Mat.device_batch(N, H, W, 1)
x4 = R.Mat.device_batch(N, H, W, 4)
run("cvtColor BGR->Gray, 4K x32 [k_px16_vec]", lambda: I.cvt_color_batch(bgr, g1, I.COLOR_BGR2GRAY), 4 * PX)
run("cvtColor RGB<->BGR, 4K x32 [k_px16_vec]", lambda: I.cvt_color_batch(bgr, out3, I.COLOR_RGB2BGR), 6 * PX)
run("cvtColor BGR->XRGB32, 4K x32 [k_px16_vec]", lambda: I.cvt_color_batch(bgr, x4, I.COLOR_BGR2XRGB32), 7 * PX)
run("cvtColor BGRA->BGR, 4K x32 [k_px16_vec]", lambda: I.cvt_color_batch(x4, out3, I.COLOR_BGRA2BGR), 7 * PX)
yuyv = batch(N, H, W, 2, seed=13)
run("cvtColor YUYV->BGR, 4K x32 [k_yuv422_vec<0>]", lambda: I.cvt_color_batch(yuyv, out3, I.COLOR_YUYV2BGR), 5 * PX)
run("cvtColor UYVY->BGR, 4K x32 [k_yuv422_vec<1>]", lambda: I.cvt_color_batch(yuyv, out3, I.COLOR_UYVY2BGR), 5 * PX)
run("cvtColor YUYV->Gray, 4K x32 [k_yuv422_vec<6>]", lambda: I.cvt_color_batch(yuyv, g1, I.COLOR_YUYV2GRAY), 3 * PX)
run("chain YUYV->BGR->GaussianBlur 5x5, 4K x32 [k_strip<YuyvGauss5Op>]", lambda: I.yuyv_to_bgr_gaussian5_batch(yuyv, out3), 5 * PX)
k3 = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)


run("filter2D 3x3, 4K BGR u8 x32 [k_strip<Filter2dU8Op<3,3>>]", lambda: I.filter2d_batch(bgr, out3, k3), 6 * PX)
run("filter2D 5x5, 4K BGR u8 x32 [k_strip<Filter2dU8Op<3,5>>]", lambda: I.filter2d_batch(bgr, out3, np.ones((5, 5), np.float32) / 25), 6 * PX)
run("filter2D 5x5, 4K BGR u8, one frame per call [k_strip<Filter2dU8Op<3,5>>]", lambda: [I.filter2d(bgr[i], out3[i], np.ones((5, 5), np.float32) / 25) for i in range(8)], 6 * 8 * H * W)
run("filter2D 7x7, 4K BGR u8 x32 [k_strip<Filter2dU8Op<3,7>>]", lambda: I.filter2d_batch(bgr, out3, np.ones((7, 7), np.float32) / 49), 6 * PX)
I.set_option("f2d.force_generic", 1)
run("filter2D 7x7, 4K BGR u8 x32, general kernel [k_filter2d<u8,7>]", lambda: I.filter2d_batch(bgr, out3, np.ones((7, 7), np.float32) / 49), 6 * PX)
I.set_option("f2d.force_generic", 0)
for b in (x4, yuyv, g1):
    b.free()
# resize
for name, dr, dc, rows_used in (("4K->1080p (exact 2x) [k_resize2x_u8<3,8>]", 1080, 1920, H), ("4K->720p (3x) [k_resize_u8w<3>]", 720, 1280, 2 * 720),
                                ("4K->1600x900 (2.4x) [k_resize_u8w<3>]", 900, 1600, 2 * 900)):
    d = R.Mat.device_batch(N, dr, dc, 3)
    run(f"resize {name}, BGR u8 x32", lambda: I.resize_batch(bgr, d), N * (rows_used * W * 3 + dr * dc * 3), "bytes = source rows touched + destination")
    d.free()
gsrc = batch(N, H, W, 1, seed=7)
for name, dr, dc in (("4K->720p (3x)", 720, 1280), ("4K->1600x900 (2.4x)", 900, 1600)):
    d = R.Mat.device_batch(N, dr, dc, 1)
    run(f"resize {name} [k_resize_u8w<1>], gray u8 x32", lambda: I.resize_batch(gsrc, d), N * (2 * dr * W + dr * dc), "bytes = source rows touched + destination")
    d.free()
gsrc.free()
d8 = R.Mat.device_batch(8, 4320, 7680, 3)
run("resize 4K->8K (exact 2x up) [k_resize_up2x_u8<3,4>], BGR u8 x8", lambda: I.resize_batch(bgr.mats[:8], d8), 8 * (H * W * 3 + 4320 * 7680 * 3))
d8.free()
M = I.get_rotation_matrix_2d(((W - 1) / 2, (H - 1) / 2), 15.0)
run("warpAffine 15 deg, 4K BGR u8 x32 [k_warp_tile<u8,3,48>]", lambda: I.warp_affine_batch(bgr, out3, M), 6 * PX)
bgr.free(); out3.free()
# f32
NF, FH, FW = 128, 1080, 1920
f = batch(NF, FH, FW, 1, R.F32, seed=3)
fo = R.Mat.device_batch(NF, FH, FW, 1, R.F32)
run("Sobel 3x3 + magnitude, 1080p f32 x128 [k_strip<Sobel3Op<0>>]", lambda: I.sobel_mag_batch(f, fo), 8 * NF * FH * FW)
for ks in (3, 5, 7):
    run(f"GaussianBlur {ks}x{ks} sigma 1.2, 1080p gray f32 x128 [k_strip<SepF32Op<{ks}>>]", lambda: I.gaussian_blur_batch(f, fo, (ks, ks), 1.2, 1.2), 8 * NF * FH * FW)
for ks, sg in ((9, 1.0), (11, 1.3), (15, 1.8)):
    run(f"GaussianBlur {ks}x{ks} sigma {sg}, 1080p gray f32 x128 [k_strip<SepF32WideOp<{ks},1>>]", lambda: I.gaussian_blur_batch(f, fo, (ks, ks), sg, sg), 8 * NF * FH * FW)
f.free(); fo.free()
f3 = batch(32, FH, FW, 3, R.F32, seed=3)
fo3 = R.Mat.device_batch(32, FH, FW, 3, R.F32)
for ks in (3, 5, 7):
    run(f"GaussianBlur {ks}x{ks} sigma 1.2, 1080p BGR f32 x32 [k_strip<SepF32CnOp<{ks},3>>]", lambda: I.gaussian_blur_batch(f3, fo3, (ks, ks), 1.2, 1.2), 24 * 32 * FH * FW)
for ks, sg in ((9, 1.0), (11, 1.3), (15, 1.8)):
    run(f"GaussianBlur {ks}x{ks} sigma {sg}, 1080p BGR f32 x32 [k_strip<SepF32WideOp<{ks},3>>]", lambda: I.gaussian_blur_batch(f3, fo3, (ks, ks), sg, sg), 24 * 32 * FH * FW)
I.set_option("sepf32.no_wide", 1)
run("GaussianBlur 11x11 sigma 2, 1080p BGR f32 x32, general kernel [k_sepfilter<f32,11>]", lambda: I.gaussian_blur_batch(f3, fo3, (11, 11), 2.0, 2.0), 24 * 32 * FH * FW)
I.set_option("sepf32.no_wide", 0)


run("filter2D 3x3, 1080p BGR f32 x32 [k_strip<Filter2dF32CnOp<3,3>>]", lambda: I.filter2d_batch(f3, fo3, k3), 24 * 32 * FH * FW)
run("filter2D 5x5, 1080p BGR f32 x32 [k_strip<Filter2dF32CnOp<5,3>>]", lambda: I.filter2d_batch(f3, fo3, np.ones((5, 5), np.float32) / 25), 24 * 32 * FH * FW)
run("filter2D 7x7, 1080p BGR f32 x32 [k_strip<Filter2dF32CnOp<7,3>>]", lambda: I.filter2d_batch(f3, fo3, np.ones((7, 7), np.float32) / 49), 24 * 32 * FH * FW)
run("filter2D 5x5, 1080p BGR f32, one frame per call [k_strip<Filter2dF32CnOp<5,3>>]", lambda: [I.filter2d(f3[i], fo3[i], np.ones((5, 5), np.float32) / 25) for i in range(8)], 24 * 8 * FH * FW)
f3.free(); fo3.free()
y8 = batch(32, FH, FW, 2, seed=6)
mg = R.Mat.device_batch(32, FH, FW, 1, R.F32)
run("chain YUYV->BGR->Gray->f32->Sobel magnitude, 1080p x32 [k_strip<YuyvSobelOp>]", lambda: I.yuyv_to_sobel_mag_batch(y8, mg), 6 * 32 * FH * FW)
y8.free(); mg.free()
S = 4096
wf = batch(16, S, S, 1, R.F32, seed=5)
wo = R.Mat.device_batch(16, S, S, 1, R.F32)
M2 = I.get_rotation_matrix_2d(((S - 1) / 2, (S - 1) / 2), 15.0)
run("warpAffine 15 deg, 4096x4096 f32 x16 [k_warp_tile<float,1,64>]", lambda: I.warp_affine_batch(wf, wo, M2), int(7.6 * 16 * S * S))
wf.free(); wo.free()
s8 = batch(16, 4320, 7680, 3, seed=4)
d4 = R.Mat.device_batch(16, 1080, 1920, 3)
run("resize 8K->1080p (exact 4x) [k_resize4x_u8c3], BGR u8 x16", lambda: I.resize_batch(s8, d4), 16 * 1080 * 1920 * 27, "27 B per destination pixel = the DRAM sector floor (15 B = the contract figure)")
sampler.stop()
