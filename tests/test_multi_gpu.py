"""GPU suite, part 2: the host side of the boundary added for SURVEY.md section 8(e) and the north star's
"pinned-host staging" -- multi-GPU batches from ONE calling thread, the coefficient broadcast, pageable Mats
through the bounce ring, caller-owned buffers page-locked in place.  Everything is compared with the CPU oracle.

The multi-GPU calls run with however many GPUs the box has (one worker thread per GPU even on a 1-GPU box, so the
fan-out code is always exercised); the cases that need two GPUs skip themselves otherwise.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from rustcv_b200 import _ffi as F

    n = C.c_int()
    F.check(F.lib.rcv_device_count(C.byref(n)))
    return n.value


@pytest.fixture(scope="module")
def multi(rcv):
    rcv.imgproc.init_multi(0)
    return rcv


def _frames(oracle, n, rows, cols, cn, seed=100):
    return [oracle.fill_u8(seed + j, rows * cols * cn).reshape(rows, cols, cn) for j in range(n)]


# ---- multi-GPU batches --------------------------------------------------------------------------------------
@pytest.mark.parametrize("where", ["host", "pinned"])
def test_gaussian_batch_multi_host_mats(multi, oracle, where):
    R = multi
    n, rows, cols = 11, 157, 203  # n not a multiple of any GPU count; odd geometry
    imgs = _frames(oracle, n, rows, cols, 3)
    if where == "host":
        srcs = [R.Mat.from_numpy(a) for a in imgs]
        dsts = [R.Mat.new(rows, cols, 3) for _ in range(n)]
    else:
        srcs, dsts = [], []
        for j, a in enumerate(imgs):
            m = R.Mat.pinned(rows, cols, 3, device=j % _ngpus())
            m.data[:] = a.ravel()
            srcs.append(m)
            dsts.append(R.Mat.pinned(rows, cols, 3, device=j % _ngpus()))
    R.imgproc.gaussian_blur_batch_multi(srcs, dsts, 0, (5, 5), 0.0, 0.0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.gaussian_blur(imgs[j], (5, 5))).all(), f"frame {j}"
    # an explicit GPU count, and a second call on the same buffers (staging ring reuse)
    R.imgproc.gaussian_blur_batch_multi(srcs, dsts, 1, (3, 3), 0.0, 0.0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.gaussian_blur(imgs[j], (3, 3))).all(), f"frame {j} (ngpus=1)"


def test_batch_multi_device_mats_follow_their_gpu(multi, oracle):
    """Device Mats run on the GPU that owns them, whatever their index in the batch."""
    R = multi
    g = _ngpus()
    n, rows, cols = 6, 96, 160
    imgs = _frames(oracle, n, rows, cols, 3, seed=300)
    owner = [(j * 5 + 1) % g for j in range(n)]  # not j mod N
    srcs = [R.Mat.from_numpy(imgs[j]).upload(owner[j]) for j in range(n)]
    dsts = [R.Mat.device_new(rows, cols, 3, device=owner[j]) for j in range(n)]
    assert [m.device for m in srcs] == owner
    R.imgproc.gaussian_blur_batch_multi(srcs, dsts, 0, (5, 5), 0.0, 0.0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.gaussian_blur(imgs[j], (5, 5))).all(), f"frame {j} on GPU {owner[j]}"


def test_every_multi_entry_point(multi, oracle):
    R = multi
    n, rows, cols = 5, 64, 96
    imgs = _frames(oracle, n, rows, cols, 3, seed=400)
    srcs = [R.Mat.from_numpy(a) for a in imgs]
    # cvtColor
    dsts = [R.Mat.new(rows, cols, 1) for _ in range(n)]
    R.imgproc.cvt_color_batch_multi(srcs, dsts, R.imgproc.COLOR_BGR2GRAY, 0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.bgr_to_gray(imgs[j])).all()
    # resize
    dsts = [R.Mat.new(40, 50, 3) for _ in range(n)]
    R.imgproc.resize_batch_multi(srcs, dsts, 0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.resize_bilinear(imgs[j], 40, 50)).all()
    # warpAffine (u8)
    M = R.imgproc.get_rotation_matrix_2d(((cols - 1) / 2, (rows - 1) / 2), 15.0)
    dsts = [R.Mat.new(rows, cols, 3) for _ in range(n)]
    R.imgproc.warp_affine_batch_multi(srcs, dsts, M, 0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.warp_affine(imgs[j], M)).all()
    # Sobel magnitude (f32)
    fimgs = [oracle.fill_f32(500 + j, rows * cols).reshape(rows, cols) for j in range(n)]
    fs = [R.Mat.from_numpy(a) for a in fimgs]
    mags = [R.Mat.new(rows, cols, 1, R.F32) for _ in range(n)]
    R.imgproc.sobel_mag_batch_multi(fs, mags, 0)
    for j in range(n):
        want = oracle.sobel3(fimgs[j])["mag"]
        d = np.abs(mags[j].to_numpy().view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert d.max() <= 1  # north_star: within 1 ULP
    # the fused chains
    yuyv = [oracle.fill_u8(600 + j, rows * cols * 2).reshape(rows, cols, 2) for j in range(n)]
    ys = [R.Mat.from_numpy(a) for a in yuyv]
    dsts = [R.Mat.new(rows, cols, 3) for _ in range(n)]
    R.imgproc.yuyv_to_bgr_gaussian5_batch_multi(ys, dsts, 0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.gaussian_blur(oracle.yuyv_to_bgr(yuyv[j]), (5, 5))).all()
    mags = [R.Mat.new(rows, cols, 1, R.F32) for _ in range(n)]
    R.imgproc.yuyv_to_sobel_mag_batch_multi(ys, mags, 0)
    for j in range(n):
        gray = oracle.convert_to(oracle.yuyv_to_gray(yuyv[j]), np.float32)
        assert (mags[j].to_numpy() == oracle.sobel3(gray)["mag"]).all()


def test_multi_rejects_frames_split_across_gpus(multi):
    R = multi
    from rustcv_b200 import _ffi as F

    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    s = R.Mat.device_new(16, 16, 3, device=0)
    d = R.Mat.device_new(16, 16, 3, device=1)
    with pytest.raises(F.RcvError) as e:
        R.imgproc.gaussian_blur_batch_multi([s], [d], 0)
    assert e.value.code == F.RCV_ERR_ARG and "no inter-GPU traffic" in str(e.value)


def test_multi_more_gpus_than_initialised(multi):
    from rustcv_b200 import _ffi as F

    R = multi
    s = R.Mat.new(16, 16, 3)
    d = R.Mat.new(16, 16, 3)
    with pytest.raises(F.RcvError) as e:
        R.imgproc.gaussian_blur_batch_multi([s], [d], _ngpus() + 1)
    assert e.value.code == F.RCV_ERR_ARG


# ---- the single collective ------------------------------------------------------------------------------------
def test_kernel_broadcast_is_consumed_by_every_gpu(multi, oracle):
    """Taps set up on the root GPU reach every GPU's bank (NCCL when there are >= 2) and the NULL-taps filter
    call launches with them: the result equals the oracle's filter with those taps on every frame."""
    R = multi
    g = _ngpus()
    taps = oracle.gaussian_kernel_q8(5, 1.3)
    got = R.imgproc.set_kernel_broadcast(np.concatenate([taps, taps]).astype(np.float32), root_device=g - 1)
    assert got.shape[0] >= g
    for d in range(g):
        assert got[d].astype(np.int32).tolist() == taps.tolist() * 2, f"GPU {d} received {got[d]}"
    n, rows, cols = 2 * g + 1, 80, 120
    imgs = _frames(oracle, n, rows, cols, 3, seed=700)
    srcs = [R.Mat.from_numpy(a) for a in imgs]
    dsts = [R.Mat.new(rows, cols, 3) for _ in range(n)]
    R.imgproc.sep_filter2d_q8_batch_multi(srcs, dsts, 0, None, None, 5, 5)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.sepfilter_u8_q8(imgs[j], taps, taps)).all(), f"frame {j}"
    # the binomial taps route to the metric kernel and give cv::GaussianBlur(5x5, sigma 0)
    R.imgproc.set_kernel_broadcast(np.array([16, 64, 96, 64, 16] * 2, np.float32), root_device=0)
    R.imgproc.sep_filter2d_q8_batch_multi(srcs, dsts, 0, None, None, 5, 5)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.gaussian_blur(imgs[j], (5, 5))).all(), f"frame {j} (binomial)"


def test_sep_filter_q8_routes_to_the_strip_ops(rcv, oracle):
    """rcv_sep_filter2d_q8 with Gaussian-like taps takes the TMA strip kernels, with anything else the general
    kernel: same numbers either way."""
    R = rcv
    img = oracle.fill_u8(801, 120 * 176 * 3).reshape(120, 176, 3)
    for taps in ([16, 64, 96, 64, 16], [64, 128, 64], [10, 60, 116, 60, 10], [2, 22, 62, 84, 62, 22, 2],
                 [0, 64, 128, 64, 0], [40, 60, 56, 60, 40], [-16, 80, 128, 80, -16], [1, 2, 3, 4, 236, 4, 3, 2, 1]):
        k = np.array(taps, np.int32)
        for where in ("host", "device"):
            s = R.Mat.from_numpy(img) if where == "host" else R.Mat.from_numpy(img).upload()
            d = R.Mat.empty() if where == "host" else s.like()
            R.imgproc.sep_filter2d(s, d, k, k)
            assert (d.to_numpy() == oracle.sepfilter_u8_q8(img, k, k)).all(), (taps, where)


# ---- pageable Mats: the bounce ring ----------------------------------------------------------------------------
def test_pageable_single_frame_banded_bounce(rcv, oracle):
    """A plain host Mat big enough to be banded: CPU copy -> H2D -> kernel -> D2H -> CPU copy, band by band."""
    R = rcv
    rows, cols = 1080, 1920
    img = oracle.fill_u8(901, rows * cols * 3).reshape(rows, cols, 3)
    want = oracle.gaussian_blur(img, (5, 5))
    for trial in range(3):  # fresh destination buffers every time (nothing is registered)
        s = R.Mat.from_numpy(img)
        d = R.Mat.empty()
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        assert (d.to_numpy() == want).all(), f"trial {trial}"
    # smaller bands (many events), a padded source and an op that leaves a column untouched
    R.imgproc.set_option("host.bounce_band_bytes", 256 << 10)
    try:
        s = R.Mat.from_numpy_strided(img, cols * 3 + 37)
        d = R.Mat.empty()
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        assert (d.to_numpy() == want).all()
        yuyv = oracle.fill_u8(902, rows * 1919 * 2).reshape(rows, 1919, 2)  # odd width: last column not written
        sy = R.Mat.from_numpy(yuyv)
        dy = R.Mat.new(rows, 1919, 3)
        dy.data[:] = 0x5A
        R.imgproc.cvt_color(sy, dy, R.imgproc.COLOR_YUYV2BGR)
        wanty = oracle.yuyv_to_bgr(yuyv)
        got = dy.to_numpy()
        assert (got[:, :1918] == wanty[:, :1918]).all() and (got[:, 1918] == 0x5A).all()
    finally:
        R.imgproc.set_option("host.bounce_band_bytes", 4 << 20)


def test_pageable_batch_and_unbandable_ops(rcv, oracle):
    R = rcv
    n, rows, cols = 9, 240, 320  # more frames than ring slots
    imgs = _frames(oracle, n, rows, cols, 3, seed=950)
    srcs = [R.Mat.from_numpy(a) for a in imgs]
    dsts = [R.Mat.new(rows, cols, 3) for _ in range(n)]
    R.imgproc.gaussian_blur_batch(srcs, dsts, (5, 5), 0.0, 0.0)
    for j in range(n):
        assert (dsts[j].to_numpy() == oracle.gaussian_blur(imgs[j], (5, 5))).all(), f"frame {j}"
    big = oracle.fill_u8(960, 1200 * 1600 * 3).reshape(1200, 1600, 3)
    d = R.Mat.empty()
    R.imgproc.resize(R.Mat.from_numpy(big), d, (400, 300))
    assert (d.to_numpy() == oracle.resize_bilinear(big, 300, 400)).all()
    # mixed: pageable source, device destination and back
    dev = R.Mat.device_new(1200, 1600, 3)
    R.imgproc.gaussian_blur(R.Mat.from_numpy(big), dev, (5, 5), 0.0)
    back = R.Mat.empty()
    R.imgproc.gaussian_blur(dev, back, (3, 3), 0.0)
    assert (back.to_numpy() == oracle.gaussian_blur(oracle.gaussian_blur(big, (5, 5)), (3, 3))).all()


# ---- caller-owned buffers page-locked in place ---------------------------------------------------------------
def test_registered_host_mats_reused_across_frames(rcv, oracle):
    """The reference reuses one Vec<u8> per Mat (videoio/mod.rs:192-199): register once, call many times."""
    R = rcv
    from rustcv_b200 import _ffi as F

    rows, cols = 1080, 1920
    s = R.Mat.new(rows, cols, 3).register()
    d = R.Mat.new(rows, cols, 3).register()
    for j in range(3):
        img = oracle.fill_u8(1000 + j, rows * cols * 3).reshape(rows, cols, 3)
        s.data[:] = img.ravel()
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        assert (d.to_numpy() == oracle.gaussian_blur(img, (5, 5))).all(), f"frame {j}"
    # a sub-range of a registered buffer is still recognised; registering twice is a no-op
    F.check(F.lib.rcv_host_register(s.data.ctypes.data, s.data.size))
    # ensure_size to another length re-registers the new buffer and releases the old one
    old = s._registered
    s.ensure_size(rows // 2, cols, 3)
    assert s._registered is not None and s._registered != old
    s.unregister()
    d.unregister()
    F.check(F.lib.rcv_host_unregister(d.data.ctypes.data))  # idempotent
    # library-owned pinned storage is not the caller's to unregister
    p = R.Mat.pinned(8, 8, 3)
    assert F.lib.rcv_host_unregister(p.data.ctypes.data) == F.RCV_ERR_ARG
    assert F.lib.rcv_pinned_free(s.data.ctypes.data) == F.RCV_ERR_ARG  # and vice versa


def test_auto_register_option(rcv, oracle):
    R = rcv
    rows, cols = 720, 1280
    img = oracle.fill_u8(1100, rows * cols * 3).reshape(rows, cols, 3)
    s = R.Mat.from_numpy(img)
    d = R.Mat.new(rows, cols, 3)
    R.imgproc.set_option("host.auto_register", 1)
    try:
        for _ in range(2):
            R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
            assert (d.to_numpy() == oracle.gaussian_blur(img, (5, 5))).all()
    finally:
        R.imgproc.set_option("host.auto_register", 0)
        from rustcv_b200 import _ffi as F

        F.check(F.lib.rcv_host_unregister(s.data.ctypes.data))
        F.check(F.lib.rcv_host_unregister(d.data.ctypes.data))


# ---- storage bookkeeping ------------------------------------------------------------------------------------------
def test_free_of_a_mat_carved_from_a_batch_is_an_argument_error(rcv):
    R = rcv
    from rustcv_b200 import _ffi as F

    b = R.Mat.device_batch(3, 16, 16, 3)
    c1 = b[1].c()
    assert F.lib.rcv_mat_free_device(C.byref(c1)) == F.RCV_ERR_ARG
    assert b"carved from a batch" in F.lib.rcv_last_error()
    b.free()


def test_rcv_device_env_selects_the_gpu():
    code = ("import ctypes as C, rustcv_b200 as R\n"
            "from rustcv_b200 import _ffi as F\n"
            "F.check(F.lib.rcv_init(-1))\n"
            "m = R.Mat.device_new(8, 8, 3)\n"
            "print('device', m.device)\n")
    n = _ngpus()
    want = n - 1
    env = dict(os.environ, RCV_DEVICE=str(want), PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    assert f"device {want}" in out.stdout


# ---- a library, not a script: concurrent callers, lifecycle ----------------------------------------------------------
def test_concurrent_callers_on_distinct_mats(rcv, oracle):
    """The header's thread contract: re-entrant and thread-safe for distinct Mats (pageable, registered and device
    Mats, single calls and batches, from several threads at once)."""
    import threading

    R = rcv
    rows, cols = 300, 420
    errs = []

    def worker(seed, kind):
        try:
            for it in range(4):
                img = oracle.fill_u8(seed * 10 + it, rows * cols * 3).reshape(rows, cols, 3)
                want = oracle.gaussian_blur(img, (5, 5))
                if kind == "host":
                    d = R.Mat.empty()
                    R.imgproc.gaussian_blur(R.Mat.from_numpy(img), d, (5, 5), 0.0)
                elif kind == "registered":
                    s = R.Mat.from_numpy(img).register()
                    d = R.Mat.new(rows, cols, 3).register()
                    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
                    s.unregister(); d.unregister()
                elif kind == "device":
                    s = R.Mat.from_numpy(img).upload()
                    d = s.like()
                    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
                else:  # a multi-GPU batch of three copies
                    srcs = [R.Mat.from_numpy(img) for _ in range(3)]
                    dsts = [R.Mat.new(rows, cols, 3) for _ in range(3)]
                    R.imgproc.gaussian_blur_batch_multi(srcs, dsts, 0, (5, 5), 0.0, 0.0)
                    d = dsts[2]
                if not (d.to_numpy() == want).all():
                    errs.append(f"{kind} seed {seed} iteration {it}: mismatch")
        except Exception as e:  # noqa: BLE001
            errs.append(f"{kind}: {type(e).__name__}: {e}")

    R.imgproc.init_multi(0)
    threads = [threading.Thread(target=worker, args=(i, k)) for i, k in enumerate(["host", "registered", "device", "multi", "host", "device"])]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errs, errs


def test_shutdown_and_reinit_in_a_fresh_process():
    """rcv_shutdown joins the worker / drain threads, frees contexts and registrations; the library can be
    initialised again afterwards (run in a subprocess: the session fixture keeps its own context)."""
    code = ("import numpy as np, rustcv_b200 as R\n"
            "from oracle import pyoracle as O\n"
            "img = O.fill_u8(7, 200 * 300 * 3).reshape(200, 300, 3)\n"
            "want = O.gaussian_blur(img, (5, 5))\n"
            "for cycle in range(3):\n"
            "    R.imgproc.init_multi(0)\n"
            "    s = R.Mat.from_numpy(img).register()\n"
            "    d = R.Mat.empty()\n"
            "    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)\n"
            "    assert (d.to_numpy() == want).all()\n"
            "    ds = [R.Mat.new(200, 300, 3) for _ in range(4)]\n"
            "    R.imgproc.gaussian_blur_batch_multi([s] * 4, ds, 0, (5, 5), 0.0, 0.0)\n"
            "    assert all((x.to_numpy() == want).all() for x in ds)\n"
            "    s._registered = None  # shutdown drops caller registrations itself\n"
            "    R.imgproc.shutdown()\n"
            "print('cycles ok')\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0 and "cycles ok" in out.stdout, out.stdout + out.stderr
