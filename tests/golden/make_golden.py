"""Generates tests/golden/cv2_golden.npz: OpenCV 4.13 outputs on SplitMix64 inputs.

The reference (RustCV @07b07dd) has no GaussianBlur / resize / BGR2GRAY / Sobel /
warpAffine / filter2D, and advertises OpenCV parity (README.md:19,30), so OpenCV's
outputs are the external pin for the oracle's specs of those ops.  Run in the build
container (cv2 4.13.0 is installed there); the GPU box only reads the .npz.

    python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

H, W = 61, 83
out = {"cv2_version": np.array(cv2.__version__), "shape": np.array([H, W])}

bgr = O.fill_u8(7, H * W * 3).reshape(H, W, 3)     # seed 7
gray = O.fill_u8(8, H * W).reshape(H, W)           # seed 8
f32 = O.fill_f32(9, H * W).reshape(H, W)           # seed 9
bgra = O.fill_u8(10, H * W * 4).reshape(H, W, 4)   # seed 10

# bit-exact u8 pins
for (kw, kh, sg) in [(3, 3, 0), (5, 5, 0), (7, 7, 0), (5, 5, 1.0), (7, 7, 1.5), (9, 9, 2.0), (0, 0, 1.2), (5, 3, 0.8)]:
    out[f"gauss_bgr_{kw}_{kh}_{sg}"] = cv2.GaussianBlur(bgr, (kw, kh), sg)
out["gauss_gray_5_5_0"] = cv2.GaussianBlur(gray, (5, 5), 0)
out["gauss_bgra_5_5_0"] = cv2.GaussianBlur(bgra, (5, 5), 0)
out["bgr2gray"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
for (dr, dc) in [(37, 64), (122, 166), (30, 41), (15, 20)]:
    out[f"resize_bgr_{dr}_{dc}"] = cv2.resize(bgr, (dc, dr), interpolation=cv2.INTER_LINEAR)
big = O.fill_u8(11, 64 * 96 * 3).reshape(64, 96, 3)  # seed 11, exact 4x case
out["resize4x_bgr"] = cv2.resize(big, (24, 16), interpolation=cv2.INTER_LINEAR)
# tolerance pins (f32: OpenCV sums in a different order / quantises warp coordinates)
gx = cv2.Sobel(f32, cv2.CV_32F, 1, 0, ksize=3)
gy = cv2.Sobel(f32, cv2.CV_32F, 0, 1, ksize=3)
out["sobel_gx"], out["sobel_gy"], out["sobel_mag"] = gx, gy, cv2.magnitude(gx, gy)
M = cv2.getRotationMatrix2D(((W - 1) / 2, (H - 1) / 2), 15.0, 1.0)
out["rotM"] = M
out["warp_f32"] = cv2.warpAffine(f32, M, (W, H), flags=cv2.INTER_LINEAR)
out["resize_f32_30_41"] = cv2.resize(f32, (41, 30), interpolation=cv2.INTER_LINEAR)
out["gauss_f32_5_5_1.1"] = cv2.GaussianBlur(f32, (5, 5), 1.1)
lap = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
out["filter2d_f32_lap"] = cv2.filter2D(f32, -1, lap)
out["filter2d_bgr"] = cv2.filter2D(bgr, -1, lap / 3 + 0.2)

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_golden.npz"), **out)
print("wrote cv2_golden.npz with", len(out), "entries")
