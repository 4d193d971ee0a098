"""One synchronous GaussianBlur call on plain (pageable) 4K host Mats through the bounce ring: band size x copy
threads sweep.  GPU box only."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:  # child: one (threads) setting, all band sizes
    import rustcv_b200 as R
    from oracle import pyoracle as O

    threads = int(sys.argv[1])
    R.imgproc.init(0)
    R.imgproc.set_option("host.copy_threads", threads)
    img = O.fill_u8(2, 2160 * 3840 * 3).reshape(2160, 3840, 3)
    s, d = R.Mat.from_numpy(img), R.Mat.new(2160, 3840, 3)
    for band in (1 << 20, 3 << 19, 2 << 20, 3 << 20, 4 << 20, 6 << 20):
        R.imgproc.set_option("host.bounce_band_bytes", band)
        for _ in range(3):
            R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        t0 = time.perf_counter()
        for _ in range(20):
            R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        ms = (time.perf_counter() - t0) / 20 * 1e3
        ok = O.crc32(d.to_numpy()) == 0x827081C8
        print(f"threads {threads:2d} band {band / (1 << 20):4.1f} MB: {ms:.3f} ms per call  parity {ok}", flush=True)
    sys.exit(0)

for t in (2, 4, 6, 8, 12):
    subprocess.run([sys.executable, __file__, str(t)], check=False)
