"""Device-resident timing of BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py).
One JSON line per config: Mpix/s, algorithmic GB/s, fraction of the measured HBM peak, and (N=1, rank 0,
CPU=1 in the environment) the CPU oracle timed beside it at 1 thread and at all host threads.
GPU box only:  python scripts/bench_configs.py [cfg ...]
               python -m torch.distributed.run --nproc-per-node N ... scripts/bench_configs.py cfg4full cfg5full
(cfg4full / cfg5full use BASELINE.json's batch sizes -- 256 / 64 frames -- sharded frame j -> rank j mod N.)
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

import time  # noqa: E402

import torch.distributed as dist  # noqa: E402

PEAK = 6549.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
R.imgproc.init(LOCAL)
stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(LOCAL), device=torch.device("cuda", LOCAL))
R.imgproc.set_blocking(False)
CPU = os.environ.get("CPU", "0") == "1" and WORLD == 1
for kv in filter(None, os.environ.get("OPT", "").split(",")):  # OPT=name:value,name:value sets library options
    R.imgproc.set_option(kv.split(":")[0], int(kv.split(":")[1]))


def cpu_time(fn, reps):
    """seconds per call of an oracle function, 1 thread and all host threads"""
    out = {}
    for label, nt in (("1t", 1), ("nt", len(os.sched_getaffinity(0)))):
        O.set_threads(nt)
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        out[label] = (time.perf_counter() - t0) / reps
    O.set_threads(1)
    out["threads"] = len(os.sched_getaffinity(0))
    return out


def fill_batch(batch, make):
    h = None
    for i in range(len(batch)):
        a = make(i)
        h = R.Mat.from_numpy(a)
        F.check(F.lib.rcv_mat_upload(C.byref(h.c()), C.byref(batch[i].c())))


def timeit(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    R.imgproc.sync(LOCAL)
    if WORLD > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    R.imgproc.sync(LOCAL)
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:  # max over ranks
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def report(name, ms, units, bytes_per_unit, extra=None):
    """units = units processed per step by THIS rank; the line reports the whole job (all ranks)."""
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    line = {"config": name, "n_gpus": WORLD, "ms_per_step": ms, "units_per_step": units * WORLD,
            "Munits_per_s": units * WORLD / ms / 1e3, "algorithmic_bytes_per_unit": bytes_per_unit,
            "achieved_GBps_per_gpu": gbs, "frac_of_measured_hbm_peak": gbs / PEAK}
    line.update(extra or {})
    if RANK == 0:
        print(json.dumps(line), flush=True)


def cfg1():
    n, h, w = 256, 480, 640
    src, dst = R.Mat.device_batch(n, h, w, 2), R.Mat.device_batch(n, h, w, 3)
    base = O.fill_u8(1, h * w * 2)
    fill_batch(src, lambda i: np.roll(base, i * 31).reshape(h, w, 2))
    ms = timeit(lambda: R.imgproc.cvt_color_batch(src, dst, R.imgproc.COLOR_YUYV2BGR))
    ok = O.crc32(dst[0].to_numpy()) == 0x0BF66518
    extra = {"crc_ok": ok}
    if CPU:  # the reference's own loop is single-threaded (videoio/mod.rs:344-371); the facade port is its restatement
        t0 = time.perf_counter()
        for _ in range(200):
            O.yuyv_to_bgr_facade(base, w, h)
        extra["cpu_ref_restated_1t_Mpix_s"] = 200 * h * w / (time.perf_counter() - t0) / 1e6
    report("cfg1 YUYV->BGR 640x480 x256", ms, n * h * w, 5, extra)
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
    extra = {"parity_frame0": ok}
    if CPU:
        img = base.reshape(h, w)
        c = cpu_time(lambda: O.sobel3(img), 8)
        extra.update({"cpu_oracle_1t_Mpix_s": h * w / c["1t"] / 1e6, "cpu_oracle_nt_Mpix_s": h * w / c["nt"] / 1e6, "cpu_threads": c["threads"]})
    report("cfg3 Sobel3x3+magnitude 1920x1080 f32 x64", ms, n * h * w, 8, extra)
    src.free(); dst.free()


def cfg4(n=16):
    n, h, w = max(1, n // WORLD), 4320, 7680
    src, dst = R.Mat.device_batch(n, h, w, 3), R.Mat.device_batch(n, h // 4, w // 4, 3)
    base = O.fill_u8(4, h * w * 3)
    fill_batch(src, lambda i: (np.roll(base, i * 31) if i else base).reshape(h, w, 3))
    ms = timeit(lambda: R.imgproc.resize_batch(src, dst), steps=10)
    ok = (O.crc32(dst[0].to_numpy()) == 0x31A84A85) if RANK == 0 else None
    extra = {"crc_ok": ok, "frames_per_gpu": n, "sector_floor_bytes_per_unit": 27, "full_source_bytes_per_unit": 51,
             "frac_vs_sector_floor": n * (h // 4) * (w // 4) * 27 / (ms * 1e-3) / 1e9 / PEAK}
    if CPU:
        img = base.reshape(h, w, 3)
        c = cpu_time(lambda: O.resize_bilinear(img, h // 4, w // 4), 4)
        px = (h // 4) * (w // 4)
        extra.update({"cpu_oracle_1t_Mpix_s": px / c["1t"] / 1e6, "cpu_oracle_nt_Mpix_s": px / c["nt"] / 1e6, "cpu_threads": c["threads"]})
    report(f"cfg4 resize 7680x4320->1920x1080 BGR u8 x{n * WORLD}", ms, n * (h // 4) * (w // 4), 15, extra)
    src.free(); dst.free()


def cfg5(n=8):
    n, s = max(1, n // WORLD), 4096
    src, dst = R.Mat.device_batch(n, s, s, 1, R.F32), R.Mat.device_batch(n, s, s, 1, R.F32)
    base = O.fill_f32(5, s * s)
    fill_batch(src, lambda i: (np.roll(base, i * 31) if i else base).reshape(s, s))
    M = R.imgproc.get_rotation_matrix_2d(((s - 1) / 2, (s - 1) / 2), 15.0)
    ms = timeit(lambda: R.imgproc.warp_affine_batch(src, dst, M), steps=10)
    extra = {"frames_per_gpu": n}
    if CPU:
        img = base.reshape(s, s)
        c = cpu_time(lambda: O.warp_affine(img, M.ravel()), 2)
        extra.update({"cpu_oracle_1t_Mpix_s": s * s / c["1t"] / 1e6, "cpu_oracle_nt_Mpix_s": s * s / c["nt"] / 1e6, "cpu_threads": c["threads"]})
    report(f"cfg5 warpAffine 15deg 4096x4096 f32 x{n * WORLD}", ms, n * s * s, 7.6, extra)
    src.free(); dst.free()


def chain(n=32):
    """raw 1080p YUYV frame -> Sobel magnitude.  The fused kernel (6 B/px) against the
    library's own chain of stand-alone kernels (YUYV2GRAY 3 B/px, convertTo 5, Sobel 8: 16 B/px; with the BGR
    intermediate a user of the separate calls would make it is 22)."""
    h, w = 1080, 1920
    src, dst = R.Mat.device_batch(n, h, w, 2), R.Mat.device_batch(n, h, w, 1, R.F32)
    base = O.fill_u8(6, h * w * 2)
    fill_batch(src, lambda i: np.roll(base, i * 31).reshape(h, w, 2))
    ms = timeit(lambda: R.imgproc.yuyv_to_sobel_mag_batch(src, dst))
    O.set_threads(8)
    f0 = base.reshape(h, w, 2)
    want = O.sobel3(O.convert_to(O.bgr_to_gray(O.yuyv_to_bgr(f0)), np.float32))["mag"]
    O.set_threads(1)
    ok = bool((dst[0].to_numpy() == want).all())
    R.imgproc.set_option("yuyvsobel.force_chain", 1)
    ms_chain = timeit(lambda: R.imgproc.yuyv_to_sobel_mag_batch(src, dst))
    R.imgproc.set_option("yuyvsobel.force_chain", 0)
    ok_chain = bool((dst[0].to_numpy() == want).all())
    report(f"chain YUYV->BGR->Gray->f32->Sobel magnitude 1920x1080 x{n}, fused kernel", ms, n * h * w, 6,
           {"parity_frame0": ok, "unfused_3_kernel_chain_ms": ms_chain, "unfused_parity_frame0": ok_chain,
            "speedup_vs_unfused": ms_chain / ms})
    src.free(); dst.free()


def chain_gauss(n=32):
    """raw 4K YUYV frame -> BGR -> GaussianBlur 5x5.  Fused kernel (5 B/px) against the
    library's two stand-alone kernels (YUYV2BGR 5 B/px + GaussianBlur 6 B/px = 11 B/px)."""
    h, w = 2160, 3840
    src, dst = R.Mat.device_batch(n, h, w, 2), R.Mat.device_batch(n, h, w, 3)
    base = O.fill_u8(7, h * w * 2)
    fill_batch(src, lambda i: np.roll(base, i * 31).reshape(h, w, 2))
    ms = timeit(lambda: R.imgproc.yuyv_to_bgr_gaussian5_batch(src, dst))
    O.set_threads(8)
    want = O.gaussian_blur(O.yuyv_to_bgr(base.reshape(h, w, 2)), (5, 5))
    O.set_threads(1)
    ok = bool((dst[0].to_numpy() == want).all())
    R.imgproc.set_option("yuyvgauss.force_chain", 1)
    ms_chain = timeit(lambda: R.imgproc.yuyv_to_bgr_gaussian5_batch(src, dst))
    R.imgproc.set_option("yuyvgauss.force_chain", 0)
    ok_chain = bool((dst[0].to_numpy() == want).all())
    report(f"chain YUYV->BGR->GaussianBlur5x5 3840x2160 x{n}, fused kernel", ms, n * h * w, 5,
           {"parity_frame0": ok, "unfused_2_kernel_chain_ms": ms_chain, "unfused_parity_frame0": ok_chain,
            "speedup_vs_unfused": ms_chain / ms})
    src.free(); dst.free()


def cvtbatch(n=32):
    """Pointwise conversions on a batch of 4K frames (one launch): the formats of SURVEY.md 8f ranks 3-4."""
    h, w = 2160, 3840
    I = R.imgproc
    bgr = R.Mat.device_batch(n, h, w, 3)
    base = O.fill_u8(12, h * w * 3)
    fill_batch(bgr, lambda i: np.roll(base, i * 31).reshape(h, w, 3))
    for name, code, dcn, bpp in (("BGR->Gray", I.COLOR_BGR2GRAY, 1, 4), ("RGB<->BGR", I.COLOR_RGB2BGR, 3, 6),
                                 ("BGR->XRGB32", I.COLOR_BGR2XRGB32, 4, 7)):
        dst = R.Mat.device_batch(n, h, w, dcn)
        ms = timeit(lambda: I.cvt_color_batch(bgr, dst, code))
        report(f"cvt {name} 3840x2160 x{n}", ms, n * h * w, bpp, {})
        dst.free()
    bgra = R.Mat.device_batch(n, h, w, 4)
    I.cvt_color_batch(bgr, bgra, I.COLOR_BGR2XRGB32)
    out3 = R.Mat.device_batch(n, h, w, 3)
    ms = timeit(lambda: I.cvt_color_batch(bgra, out3, I.COLOR_BGRA2BGR))
    report(f"cvt BGRA->BGR 3840x2160 x{n}", ms, n * h * w, 7, {})
    bgra.free()
    yuyv = R.Mat.device_batch(n, h, w, 2)
    yb = O.fill_u8(13, h * w * 2)
    fill_batch(yuyv, lambda i: np.roll(yb, i * 31).reshape(h, w, 2))
    for name, code in (("YUYV->BGR", I.COLOR_YUYV2BGR), ("UYVY->BGR", I.COLOR_UYVY2BGR)):
        ms = timeit(lambda: I.cvt_color_batch(yuyv, out3, code))
        report(f"cvt {name} 3840x2160 x{n}", ms, n * h * w, 5, {})
    g = R.Mat.device_batch(n, h, w, 1)
    ms = timeit(lambda: I.cvt_color_batch(yuyv, g, I.COLOR_YUYV2GRAY))
    report(f"cvt YUYV->Gray 3840x2160 x{n}", ms, n * h * w, 3, {})
    f = R.Mat.device_batch(n, h, w, 1, R.F32)
    ms = timeit(lambda: [I.convert_to(g[i], f[i], R.F32, 1.0 / 255) for i in range(4)])
    report("convertTo u8->f32 3840x2160 (4 launches of 1 frame)", ms, 4 * h * w, 5, {})
    for b in (bgr, out3, yuyv, g, f):
        b.free()


def resizebatch(n=16):
    """Bilinear resize on a batch of 4K BGR frames (exact 2x: fast path; the rest: general gather kernel);
    bytes = source rows touched + destination."""
    h, w = 2160, 3840
    src = R.Mat.device_batch(n, h, w, 3)
    base = O.fill_u8(14, h * w * 3)
    fill_batch(src, lambda i: np.roll(base, i * 31).reshape(h, w, 3))
    for name, dr, dc, rows_used in (("4K->1080p (2x)", 1080, 1920, h), ("4K->720p (3x)", 720, 1280, 2 * 720),
                                    ("4K->1600x900 (2.4x)", 900, 1600, 2 * 900), ("4K->8K (upscale 2x)", 4320, 7680, h)):
        dst = R.Mat.device_batch(n, dr, dc, 3)
        ms = timeit(lambda: R.imgproc.resize_batch(src, dst), steps=10)
        want = O.resize_bilinear(base.reshape(h, w, 3), dr, dc) if dr <= 1080 else None
        ok = bool((dst[0].to_numpy() == want).all()) if want is not None else None
        report(f"resize {name} BGR u8 x{n}", ms, n * dr * dc, (rows_used * w * 3 + dr * dc * 3) / (dr * dc),
               {"parity_frame0": ok})
        dst.free()
    src.free()


def cfg5sweep(n=8):
    """warpAffine tile height sweep (warp.tile_rows = 32 / 48 / 64 / automatic), parity of frame 0 each time."""
    s = 4096
    src, dst = R.Mat.device_batch(n, s, s, 1, R.F32), R.Mat.device_batch(n, s, s, 1, R.F32)
    base = O.fill_f32(5, s * s)
    fill_batch(src, lambda i: (np.roll(base, i * 31) if i else base).reshape(s, s))
    O.set_threads(len(os.sched_getaffinity(0)))
    for ang in (15.0, 90.0, 33.0):
        M = R.imgproc.get_rotation_matrix_2d(((s - 1) / 2, (s - 1) / 2), ang)
        want = O.warp_affine(base.reshape(s, s), M.ravel())
        for th in (32, 48, 64, 0):
            R.imgproc.set_option("warp.tile_rows", th)
            ms = timeit(lambda: R.imgproc.warp_affine_batch(src, dst, M), steps=10)
            got = dst[0].to_numpy()
            ok = bool((got == want).all())
            report(f"cfg5 warpAffine {ang:g}deg 4096x4096 f32 x{n} tile_rows={th}", ms, n * s * s, 7.6, {"bit_exact_frame0": ok})
    R.imgproc.set_option("warp.tile_rows", 0)
    O.set_threads(1)
    # u8 BGR 4K, 15 degrees
    h, w = 2160, 3840
    src8, dst8 = R.Mat.device_batch(n, h, w, 3), R.Mat.device_batch(n, h, w, 3)
    b8 = O.fill_u8(5, h * w * 3)
    fill_batch(src8, lambda i: np.roll(b8, i * 31).reshape(h, w, 3))
    M = R.imgproc.get_rotation_matrix_2d(((w - 1) / 2, (h - 1) / 2), 15.0)
    for th in (32, 48, 64, 0):
        R.imgproc.set_option("warp.tile_rows", th)
        ms = timeit(lambda: R.imgproc.warp_affine_batch(src8, dst8, M), steps=10)
        report(f"warpAffine 15deg 3840x2160 BGR u8 x{n} tile_rows={th}", ms, n * h * w, 6, {})
    R.imgproc.set_option("warp.tile_rows", 0)
    src.free(); dst.free(); src8.free(); dst8.free()


ALL = {"cvtbatch": cvtbatch, "resizebatch": resizebatch, "chain": chain, "chain_gauss": chain_gauss, "cfg5sweep": cfg5sweep, "cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5,
       "cfg4full": lambda: cfg4(256), "cfg5full": lambda: cfg5(64)}
for name in (sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]):
    ALL[name]()
if WORLD > 1:
    dist.barrier()
    dist.destroy_process_group()
