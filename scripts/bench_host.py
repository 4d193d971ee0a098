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
