"""Host-resident single-call latency: pageable vs pinned Mats through the synchronous C-ABI call (GPU box only)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rustcv_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

R.imgproc.init(0)
H, W = 2160, 3840
a = O.fill_u8(2, H * W * 3).reshape(H, W, 3)


def t(fn, n=10):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


src_p = R.Mat.from_numpy(a)
dst_p = R.Mat.new(H, W, 3)
src_q = R.Mat.pinned(H, W, 3)
src_q.data[:] = a.ravel()
dst_q = R.Mat.pinned(H, W, 3)
ms_p = t(lambda: R.imgproc.gaussian_blur(src_p, dst_p, (5, 5), 0.0))
ms_q = t(lambda: R.imgproc.gaussian_blur(src_q, dst_q, (5, 5), 0.0))
assert O.crc32(dst_p.to_numpy()) == 0x827081C8 and O.crc32(dst_q.to_numpy()) == 0x827081C8
mb = H * W * 3 / 1e6
print(f"single 4K GaussianBlur, host Mats: pageable {ms_p:.2f} ms ({2 * mb / ms_p / 1e3 * 1e3 / 1e3:.1f} GB/s moved), "
      f"pinned {ms_q:.2f} ms ({2 * mb / ms_q:.1f} MB/ms)")
n = 16
srcs = [R.Mat.from_numpy(a) for _ in range(n)]
dsts = [R.Mat.new(H, W, 3) for _ in range(n)]
ms_b = t(lambda: R.imgproc.gaussian_blur_batch(srcs, dsts), n=3)
print(f"batch of {n} pageable 4K frames: {ms_b:.1f} ms = {n * H * W / ms_b / 1e3:.0f} Mpix/s")

# ---- pinned Mats: whole-frame staging vs the banded pipeline (H2D / kernel / D2H of one frame overlap) ----------
srcs_q = [R.Mat.pinned(H, W, 3) for _ in range(n)]
dsts_q = [R.Mat.pinned(H, W, 3) for _ in range(n)]
for m in srcs_q:
    m.data[:] = a.ravel()
f = O.fill_f32(3, 1080 * 1920).reshape(1080, 1920)
fs, fd = R.Mat.pinned(1080, 1920, 1, R.F32), R.Mat.pinned(1080, 1920, 1, R.F32)
fs.data[:] = f.view(np.uint8).ravel()
y = O.fill_u8(1, 1080 * 1920 * 2).reshape(1080, 1920, 2)
ys, yd, ym = R.Mat.pinned(1080, 1920, 2), R.Mat.pinned(1080, 1920, 3), R.Mat.pinned(1080, 1920, 1, R.F32)
ys.data[:] = y.ravel()
want_s, want_y = O.sobel3(f)["mag"], O.yuyv_to_bgr(y)
for dw, bb in [(0, 0), (0, 6 << 20)] + [(1, b) for b in (0, 1 << 19, 1 << 20, 2 << 20, 3 << 20, 4 << 20, 6 << 20)]:
    R.imgproc.set_option("host.direct_write", dw)
    R.imgproc.set_option("host.band_bytes", bb)
    for m in dsts_q:
        m.data[:] = 0
    dst_q.data[:] = 0
    fd.data[:] = 0
    yd.data[:] = 0
    ms1 = t(lambda: R.imgproc.gaussian_blur(src_q, dst_q, (5, 5), 0.0), n=20)
    msb = t(lambda: R.imgproc.gaussian_blur_batch(srcs_q, dsts_q), n=5)
    ms_s = t(lambda: R.imgproc.sobel_mag(fs, fd), n=20)
    ms_y = t(lambda: R.imgproc.cvt_color(ys, yd, R.imgproc.COLOR_YUYV2BGR), n=20)
    ms_c = t(lambda: R.imgproc.yuyv_to_sobel_mag(ys, ym), n=20)
    ok = (O.crc32(dst_q.to_numpy()) == 0x827081C8 and all(O.crc32(m.to_numpy()) == 0x827081C8 for m in dsts_q)
          and bool((fd.to_numpy() == want_s).all() and (yd.to_numpy() == want_y).all()))
    print(f"pinned, host.direct_write={dw} host.band_bytes={bb / (1 << 20):g} MB: single 4K GaussianBlur {ms1:.3f} ms; batch of {n}: {msb:.2f} ms = "
          f"{n * H * W / msb / 1e3:.0f} Mpix/s ({n * mb / msb:.1f} GB/s each way); 1080p Sobel f32 {ms_s:.3f} ms, "
          f"YUYV->BGR {ms_y:.3f} ms, YUYV->Sobel fused {ms_c:.3f} ms; parity {ok}")
R.imgproc.set_option("host.band_bytes", 6 << 20)
R.imgproc.set_option("host.direct_write", 0)
