"""CPU suite, part 4: the oracle under AddressSanitizer + UndefinedBehaviorSanitizer.
The oracle defines most of the results the GPU path is held to, so its own memory safety is checked: an
instrumented build runs every op once on odd-sized, padded images (1-row / 1-column cases included) in a
subprocess; any report aborts it."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_oracle_is_clean_under_asan_and_ubsan(tmp_path):
    so = tmp_path / "librcv_oracle_san.so"
    cmd = ["gcc", "-O1", "-g", "-std=c99", "-fPIC", "-ffp-contract=off", "-pthread", "-fsanitize=address,undefined",
           "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-shared", "-o", str(so),
           os.path.join(ROOT, "oracle", "rcv_oracle.c"), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("asan" in r.stderr.lower() or "sanitize" in r.stderr.lower()):
        pytest.skip("this gcc has no sanitizer runtime")
    assert r.returncode == 0, r.stderr
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    code = textwrap.dedent(f"""
        import ctypes as C, sys
        import numpy as np
        sys.path.insert(0, {ROOT!r})
        from oracle import pyoracle as O
        O._lib = C.CDLL({str(so)!r})          # every wrapper below now calls the instrumented build
        O.lib = lambda: O._lib
        for (h, w) in ((1, 1), (1, 7), (9, 1), (5, 6), (37, 53)):
            u3 = O.fill_u8(1, h * w * 3).reshape(h, w, 3)
            f1 = O.fill_f32(2, h * w).reshape(h, w)
            O.gaussian_blur(u3, (5, 5)); O.gaussian_blur(u3, (3, 3)); O.gaussian_blur(u3, (11, 11), 2.0, 2.0)
            O.gaussian5_fast(u3)
            O.gaussian_blur(f1, (7, 7), 1.5, 1.5)
            O.filter2d(u3, np.ones((3, 3), np.float32) / 9, 0.5); O.filter2d(f1, np.ones((5, 3), np.float32), 0.0)
            O.sobel3(f1, ("mag", "gx", "gy"))
            O.resize_bilinear(u3, max(1, h // 2), max(1, w * 2)); O.resize_bilinear(f1, h + 3, max(1, w // 3))
            M = O.rotation_matrix((w - 1) / 2, (h - 1) / 2, 15.0)
            O.warp_affine(f1, M); O.warp_affine(u3, M)
            O.bgr_to_gray(u3); O.swap_rb(u3); O.bgr_to_xrgb32(u3); O.convert_to(u3, np.float32, 1 / 255.0, 0.0)
            O.bgra_to_bgr(O.fill_u8(3, h * w * 4).reshape(h, w, 4))
            if w % 2 == 0:
                y = O.fill_u8(4, h * w * 2).reshape(h, w, 2)
                O.yuyv_to_bgr(y); O.uyvy_to_bgr(y); O.yuyv_to_gray(y)
                O.yuyv_to_bgr_facade(y.ravel(), w, h); O.yuyv_to_bgr_camera(y.ravel(), w, h)
        print("sanitized oracle ok")
    """)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "sanitized oracle ok" in out.stdout, out.stderr[-3000:]
