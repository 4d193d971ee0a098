"""Generates tests/golden/mjpeg_golden.npz: three small JPEG frames (the bytes a MJPG camera would deliver) and
their decode by libjpeg-turbo through OpenCV 4.13 (cv2.imdecode) -- the decoder family the reference uses
(the turbojpeg crate, rustcv/src/videoio/mod.rs:205-232).  JPEG decoders are not bit-identical to one another
(IDCT rounding, chroma upsampling filters), so the GPU test compares within a stated tolerance.

    python tests/golden/make_mjpeg_golden.py        (build container; the GPU box only reads the .npz)
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(h, w, seed):
    """A camera-like frame: smooth gradients, a few sinusoids, two flat rectangles, mild sensor noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        img[..., c] = (90 + 60 * np.sin(x / (9.0 + 3 * c) + seed) * np.cos(y / (13.0 - 2 * c))
                       + 50 * (x / w) + 30 * (y / h) * (c - 1))
    img[h // 5:h // 2, w // 6:w // 3] = (40, 180, 220)
    img[h // 2:h - h // 6, w // 2:w - w // 8] = (200, 60, 30)
    img += rng.normal(0, 2.0, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


out = {"cv2_version": np.array(cv2.__version__)}
cases = [("444_q95", 64, 96, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 95),
         ("420_q90", 120, 160, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 90),
         ("422_q85", 240, 320, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, 85),
         ("420_odd_q92", 77, 121, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 92),   # odd sizes: chroma edge columns / rows
         ("422_odd_q92", 51, 99, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, 92)]
for name, h, w, sf, q in cases:
    ok, buf = cv2.imencode(".jpg", scene(h, w, len(name) + h), [cv2.IMWRITE_JPEG_QUALITY, q,
                                                                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf])
    assert ok
    out[f"jpeg_{name}"] = np.frombuffer(buf.tobytes(), np.uint8)
    out[f"bgr_{name}"] = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    print(name, h, w, len(buf), "bytes")
ok, buf = cv2.imencode(".jpg", cv2.cvtColor(scene(90, 130, 5), cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 90])
out["jpeg_gray_q90"] = np.frombuffer(buf.tobytes(), np.uint8)
out["bgr_gray_q90"] = cv2.imdecode(buf, cv2.IMREAD_COLOR)
np.savez_compressed(os.path.join(HERE, "mjpeg_golden.npz"), **out)
