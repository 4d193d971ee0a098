"""Which direction of zero-copy is slow?  4K GaussianBlur with one side in pinned host memory accessed directly
by the kernel (host.zero_copy=1) and the other side in HBM.  GPU box only."""
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
mb = H * W * 3 / 1e6


def t(fn, n=20):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


hs, hd = R.Mat.pinned(H, W, 3), R.Mat.pinned(H, W, 3)
hs.data[:] = a.ravel()
ds = R.Mat.from_numpy(a).upload()
dd = ds.like()
R.imgproc.set_option("host.zero_copy", 1)
for br in (0, 60, 124, 252):
    R.imgproc.set_option("gauss.band_rows", br)
    ms_r = t(lambda: R.imgproc.gaussian_blur(hs, dd, (5, 5), 0.0))
    ok_r = O.crc32(dd.to_numpy()) == 0x827081C8
    hd.data[:] = 0
    ms_w = t(lambda: R.imgproc.gaussian_blur(ds, hd, (5, 5), 0.0))
    ok_w = O.crc32(hd.to_numpy()) == 0x827081C8
    print(f"band_rows={br}: zero-copy READ (pinned -> HBM) {ms_r:.3f} ms = {mb / ms_r:.1f} GB/s {ok_r}; "
          f"zero-copy WRITE (HBM -> pinned) {ms_w:.3f} ms = {mb / ms_w:.1f} GB/s {ok_w}")
R.imgproc.set_option("gauss.band_rows", 0)
R.imgproc.set_option("host.zero_copy", 0)
ms_r = t(lambda: R.imgproc.gaussian_blur(hs, dd, (5, 5), 0.0))
ms_w = t(lambda: R.imgproc.gaussian_blur(ds, hd, (5, 5), 0.0))
print(f"staged: pinned -> HBM {ms_r:.3f} ms = {mb / ms_r:.1f} GB/s; HBM -> pinned {ms_w:.3f} ms = {mb / ms_w:.1f} GB/s")
