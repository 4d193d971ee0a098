"""GPU suite: parity of the CUDA path (through the C ABI) against the CPU oracle.

Bit-exact for every u8 op; f32 ops are bit-exact too where the oracle fixes the
operation order (Sobel, resize, warpAffine, separable/dense filters), with the 1-ULP
tolerance BASELINE.json's north_star allows stated in `assert_f32`.
Sizes are chosen so the oracle finishes in seconds; the full BASELINE.json sizes are
covered by CRC goldens and size-independent properties at the end.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


# ---- helpers --------------------------------------------------------------------------
def assert_same(got, want, name):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    bad = got != want
    if bad.ndim == 3:
        bad2 = bad.any(axis=2)
    else:
        bad2 = bad
    if bad.any():
        ys, xs = np.nonzero(bad2)
        first = [(int(y), int(x), got[y, x].tolist(), want[y, x].tolist()) for y, x in list(zip(ys, xs))[:6]]
        raise AssertionError(
            f"{name}: {int(bad.sum())} of {bad.size} elements differ; rows {ys.min()}..{ys.max()} "
            f"cols {xs.min()}..{xs.max()}; distinct rows {len(set(ys.tolist()))} distinct cols {len(set(xs.tolist()))}; "
            f"first (y, x, got, want): {first}")


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def assert_f32(got, want, name, max_ulp=1):
    """north_star tolerance: within 1 ULP for f32."""
    d = ulp_diff(got, want)
    if d.max() > max_ulp:
        ys, xs = np.nonzero(d > max_ulp)
        first = [(int(y), int(x), float(got[y, x]), float(want[y, x])) for y, x in list(zip(ys, xs))[:6]]
        raise AssertionError(f"{name}: max ulp {int(d.max())}, {int((d > max_ulp).sum())} elements beyond {max_ulp} ulp; "
                             f"rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; first {first}")


def mats(R, a, where, step=None):
    """A Mat holding `a` as host (packed or padded), pinned or device storage."""
    if where == "host":
        return R.Mat.from_numpy(a) if step is None else R.Mat.from_numpy_strided(a, step)
    if where == "pinned":
        cn = 1 if a.ndim == 2 else a.shape[2]
        m = R.Mat.pinned(a.shape[0], a.shape[1], cn, R.F32 if a.dtype == np.float32 else R.U8)
        m.data[:] = np.ascontiguousarray(a).view(np.uint8).ravel()
        return m
    return R.Mat.from_numpy(a).upload()


def upload_into(R, a, dev_mat):
    from rustcv_b200 import _ffi as F
    h = R.Mat.from_numpy(a)  # keep the host Mat alive across the call
    F.check(F.lib.rcv_mat_upload(C.byref(h.c()), C.byref(dev_mat.c())))


def out_like(R, src, where, channels=None, depth=None, rows=None, cols=None):
    """A zero-filled destination in the same kind of storage (ops that leave pixels untouched -- the
    odd last column of a YUYV row -- are compared against the oracle's zero-initialised output)."""
    if where == "host":
        return R.Mat.empty()
    m = src.like(channels=channels, depth=depth, rows=rows, cols=cols)
    if where == "pinned":
        m.data[:] = 0
    else:
        z = R.Mat.new(m.rows, m.cols, m.channels, m.depth)
        if m.rows and m.cols:
            from rustcv_b200 import _ffi as F
            F.check(F.lib.rcv_mat_upload(C.byref(z.c()), C.byref(m.c())))
    return m


WHERE = ["host", "device", "pinned"]


# ---- conversions (the reference's real loops) ----------------------------------------------
def test_reference_unit_tests_on_gpu(rcv):
    R = rcv
    # rustcv-camera/src/decode.rs:235-273 run against the CUDA path
    d = R.Mat.empty()
    R.imgproc.yuyv_to_bgr(R.Mat.from_numpy(np.array([[[235, 128], [235, 128]]], np.uint8)), d)
    assert (d.to_numpy() > 240).all() and (d.to_numpy() == 255).all()
    R.imgproc.yuyv_to_bgr(R.Mat.from_numpy(np.array([[[16, 128], [16, 128]]], np.uint8)), d)
    assert (d.to_numpy() < 10).all()
    R.imgproc.cvt_color(R.Mat.from_numpy(np.array([[[255, 0, 0], [0, 255, 0]]], np.uint8)), d, R.imgproc.COLOR_RGB2BGR)
    assert d.to_numpy().ravel().tolist() == [0, 0, 255, 0, 255, 0]


def test_yuyv_cfg1_packed_facade_contract(rcv, oracle):
    """BASELINE.json config 1: one 640x480 YUYV frame into a Mat, facade contract."""
    R = rcv
    src = oracle.fill_u8(1, 640 * 480 * 2)
    m = R.Mat.empty()
    assert R.videoio.decode_frame(src, 640, 480, R.videoio.YUYV, m)
    assert (m.rows, m.cols, m.channels, m.step) == (480, 640, 3, 1920)
    assert oracle.crc32(m.data) == 0x0BF66518
    assert m.data[:6].tolist() == [133, 213, 220, 0, 0, 0]
    # odd pixel count: w*h/2 macro-pixels as one packed run, last pixel untouched
    src = oracle.fill_u8(3, 5 * 3 * 2)
    m = R.Mat.empty()
    R.videoio.decode_frame(src, 5, 3, R.videoio.YUYV, m)
    st, want = oracle.yuyv_to_bgr_facade(src, 5, 3)
    assert st == 0 and (m.data == want).all()
    # BGRA packed (videoio/mod.rs:385-399)
    src = oracle.fill_u8(4, 33 * 7 * 4)
    R.videoio.decode_frame(src, 33, 7, R.videoio.BGRA, m)
    assert (m.data == oracle.bgra_to_bgr_facade(src, 33, 7)[1]).all()


@pytest.mark.parametrize("where", WHERE)
@pytest.mark.parametrize("shape", [(480, 640), (7, 2), (5, 37), (64, 1024), (3, 1)])
def test_yuv422_strided(rcv, oracle, where, shape):
    R = rcv
    h, w = shape
    src = oracle.fill_u8(21, h * w * 2).reshape(h, w, 2)
    for code, fn in ((R.imgproc.COLOR_YUYV2BGR, oracle.yuyv_to_bgr), (R.imgproc.COLOR_UYVY2BGR, oracle.uyvy_to_bgr),
                     (R.imgproc.COLOR_YUYV2GRAY, oracle.yuyv_to_gray)):
        s = mats(R, src, where)
        d = out_like(R, s, where, channels=1 if code == R.imgproc.COLOR_YUYV2GRAY else 3)
        R.imgproc.cvt_color(s, d, code)
        assert_same(d.to_numpy(), fn(src), f"cvt {code} {shape} {where}")


def test_yuv_formula_exhaustive_on_gpu(rcv):
    """All 2^24 (Y,U,V) triples through the CUDA kernel vs numpy (videoio/mod.rs:352-369)."""
    R = rcv
    u, v = np.meshgrid(np.arange(256, dtype=np.int32), np.arange(256, dtype=np.int32), indexing="ij")
    u, v = u.ravel(), v.ravel()
    frame = np.empty((256, 65536, 4), np.uint8)
    for y in range(256):
        frame[y, :, 0] = y
        frame[y, :, 2] = 255 - y
    frame[:, :, 1] = u[None, :]
    frame[:, :, 3] = v[None, :]
    d = R.Mat.empty()
    R.imgproc.yuyv_to_bgr(R.Mat.from_numpy(frame.reshape(256, 65536 * 2, 2)), d)
    got = d.to_numpy().reshape(256, 65536, 6)
    for y in (0, 1, 15, 16, 17, 100, 128, 200, 234, 235, 236, 254, 255):
        for k, yy in ((0, y), (3, 255 - y)):
            c = yy - 16
            b = np.clip((298 * c + 516 * (u - 128) + 128) >> 8, 0, 255)
            g = np.clip((298 * c - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8, 0, 255)
            r = np.clip((298 * c + 409 * (v - 128) + 128) >> 8, 0, 255)
            assert (got[y, :, k] == b).all() and (got[y, :, k + 1] == g).all() and (got[y, :, k + 2] == r).all(), y
    # every row against the vectorised formula via a checksum of all 2^24 x 2 pixels
    c0 = frame[:, :, 0].astype(np.int32) - 16
    uu, vv = frame[:, :, 1].astype(np.int32) - 128, frame[:, :, 3].astype(np.int32) - 128
    assert (got[:, :, 0] == np.clip((298 * c0 + 516 * uu + 128) >> 8, 0, 255)).all()
    assert (got[:, :, 1] == np.clip((298 * c0 - 100 * uu - 208 * vv + 128) >> 8, 0, 255)).all()
    assert (got[:, :, 2] == np.clip((298 * c0 + 409 * vv + 128) >> 8, 0, 255)).all()
    # YUYV -> Gray over the same 2^24 triples: the kernel computes two pixels per operation with the gray
    # coefficients split into bytes (cvt_math.cuh gray_pair); the definition is OpenCV's 15-bit formula on the BGR above
    dg = R.Mat.empty()
    R.imgproc.cvt_color(R.Mat.from_numpy(frame.reshape(256, 65536 * 2, 2)), dg, R.imgproc.COLOR_YUYV2GRAY)
    gray = dg.to_numpy().reshape(256, 65536, 2)
    for k in (0, 1):
        b, g, r = (got[:, :, 3 * k + i].astype(np.int32) for i in range(3))
        assert (gray[:, :, k] == ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15)).all(), k


@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("shape", [(31, 45), (16, 64), (1, 1), (9, 257)])
def test_byte_format_conversions(rcv, oracle, where, shape):
    R = rcv
    h, w = shape
    bgr = oracle.fill_u8(22, h * w * 3).reshape(h, w, 3)
    bgra = oracle.fill_u8(23, h * w * 4).reshape(h, w, 4)
    I = R.imgproc
    for code, src, fn, cn in ((I.COLOR_BGRA2BGR, bgra, oracle.bgra_to_bgr, 3), (I.COLOR_RGB2BGR, bgr, oracle.swap_rb, 3),
                              (I.COLOR_BGR2GRAY, bgr, oracle.bgr_to_gray, 1), (I.COLOR_BGR2XRGB32, bgr, None, 4)):
        s = mats(R, src, where)
        d = out_like(R, s, where, channels=cn)
        I.cvt_color(s, d, code)
        if fn is None:
            want = oracle.bgr_to_xrgb32(bgr).view(np.uint8).reshape(h, w, 4)
        else:
            want = fn(src)
        assert_same(d.to_numpy(), want, f"cvt {code} {shape} {where}")
    # padded steps (Mat.step > cols*channels), unaligned -> scalar kernels
    s = R.Mat.from_numpy_strided(bgr, step=w * 3 + 5)
    d = R.Mat.empty()
    I.cvt_color(s, d, I.COLOR_BGR2GRAY)
    assert_same(d.to_numpy(), oracle.bgr_to_gray(bgr), "gray strided")


def test_nv12(rcv, oracle):
    R = rcv
    h, w = 34, 58
    y = oracle.fill_u8(24, h * w).reshape(h, w)
    uv = oracle.fill_u8(25, (h // 2) * (w // 2) * 2).reshape(h // 2, w // 2, 2)
    d = R.Mat.empty()
    R.imgproc.nv12_to_bgr(R.Mat.from_numpy(y), R.Mat.from_numpy(uv), d)
    assert_same(d.to_numpy(), oracle.nv12_to_bgr(y, uv.reshape(h // 2, w)), "nv12")
    # widths that are multiples of 16 on aligned storage take the vector kernel (16 pixels per thread)
    for (h, w), where in (((48, 64), "device"), ((270, 480), "device"), ((34, 1600), "pinned"), ((480, 640), "host")):
        y = oracle.fill_u8(26 + w, h * w).reshape(h, w)
        y[0:4, 0:32] = 255  # saturating lumas with random chroma
        uv = oracle.fill_u8(27 + w, (h // 2) * (w // 2) * 2).reshape(h // 2, w // 2, 2)
        ym, um = mats(R, y, where), mats(R, uv, where)
        d = out_like(R, ym, where, channels=3)
        R.imgproc.nv12_to_bgr(ym, um, d)
        assert_same(d.to_numpy(), oracle.nv12_to_bgr(y, uv.reshape(h // 2, w)), f"nv12 vec {h}x{w} {where}")


# ---- GaussianBlur ---------------------------------------------------------------------------
GAUSS_SHAPES = [(61, 83, 3), (8, 8, 3), (64, 160, 3), (97, 161, 3), (300, 500, 3), (40, 1000, 3), (500, 24, 3),
                (33, 160, 1), (50, 481, 1), (45, 70, 2), (61, 83, 4), (72, 120, 4), (250, 321, 3)]


@pytest.mark.parametrize("where", WHERE)
@pytest.mark.parametrize("shape", GAUSS_SHAPES)
def test_gaussian5_binomial_strip_kernel(rcv, oracle, where, shape):
    """The metric kernel (k_strip<Gauss5Op>): every edge combination vs the oracle."""
    R = rcv
    h, w, cn = shape
    a = oracle.fill_u8(30 + h + w, h * w * cn).reshape(h, w, cn)
    if cn == 1:
        a = a.reshape(h, w)
    s = mats(R, a, where)
    d = out_like(R, s, where)
    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
    assert_same(d.to_numpy(), oracle.gaussian_blur(a, (5, 5)), f"gauss5 {shape} {where}")


@pytest.mark.parametrize("where", WHERE)
@pytest.mark.parametrize("shape", GAUSS_SHAPES)
def test_gaussian3_binomial_strip_kernel(rcv, oracle, where, shape):
    """GaussianBlur ksize 3, sigma 0 (k_strip<Gauss3Op>): every edge combination vs the oracle, and bit-identical
    to the any-sigma op it replaces for these taps."""
    R = rcv
    h, w, cn = shape
    a = oracle.fill_u8(33 + h + w, h * w * cn).reshape(h, w, cn)
    if cn == 1:
        a = a.reshape(h, w)
    s = mats(R, a, where)
    d = out_like(R, s, where)
    R.imgproc.gaussian_blur(s, d, (3, 3), 0.0)
    want = oracle.gaussian_blur(a, (3, 3))
    assert_same(d.to_numpy(), want, f"gauss3 {shape} {where}")
    if where == "device":
        for band in (8, 12, 100):
            R.imgproc.set_option("gauss.band_rows", band)
            try:
                d2 = s.like()
                R.imgproc.gaussian_blur(s, d2, (3, 3), 0.0)
                assert_same(d2.to_numpy(), want, f"gauss3 {shape} band {band}")
            finally:
                R.imgproc.set_option("gauss.band_rows", 0)
        R.imgproc.set_option("gauss.no_binomial3", 1)
        try:
            d3 = s.like()
            R.imgproc.gaussian_blur(s, d3, (3, 3), 0.0)
            assert_same(d3.to_numpy(), want, f"any-sigma op on binomial taps {shape}")
        finally:
            R.imgproc.set_option("gauss.no_binomial3", 0)


@pytest.mark.parametrize("band_rows", [8, 12, 28, 36, 100])
def test_gaussian5_band_seams(rcv, oracle, band_rows):
    """Band boundaries (warm-up rows) and multi-chunk pipelines must be seamless."""
    R = rcv
    a = oracle.fill_u8(41, 211 * 1003 * 3).reshape(211, 1003, 3)
    R.imgproc.set_option("gauss.band_rows", band_rows)
    try:
        s = mats(R, a, "device")
        d = s.like()
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        assert_same(d.to_numpy(), oracle.gaussian_blur(a, (5, 5)), f"band_rows {band_rows}")
    finally:
        R.imgproc.set_option("gauss.band_rows", 0)


def test_gaussian5_generic_kernel_agrees(rcv, oracle):
    R = rcv
    a = oracle.fill_u8(42, 97 * 161 * 3).reshape(97, 161, 3)
    R.imgproc.set_option("gauss.force_generic", 1)
    try:
        d = R.Mat.empty()
        R.imgproc.gaussian_blur(R.Mat.from_numpy(a), d, (5, 5), 0.0)
        assert_same(d.to_numpy(), oracle.gaussian_blur(a, (5, 5)), "generic 5x5")
    finally:
        R.imgproc.set_option("gauss.force_generic", 0)


@pytest.mark.parametrize("shape", [(1, 1, 3), (1, 9, 3), (9, 1, 3), (2, 2, 1), (3, 5, 3), (5, 3, 4), (7, 300, 3), (300, 7, 3)])
def test_gaussian_tiny_images(rcv, oracle, shape):
    R = rcv
    h, w, cn = shape
    a = oracle.fill_u8(43, h * w * cn).reshape(h, w, cn)
    if cn == 1:
        a = a.reshape(h, w)
    for ks in ((5, 5), (3, 3)):
        d = R.Mat.empty()
        R.imgproc.gaussian_blur(R.Mat.from_numpy(a), d, ks, 0.0)
        assert_same(d.to_numpy(), oracle.gaussian_blur(a, ks), f"tiny {shape} {ks}")


@pytest.mark.parametrize("ks,sigma", [((3, 3), 0), ((7, 7), 0), ((5, 5), 1.0), ((7, 7), 1.5), ((9, 9), 2.0), ((0, 0), 1.2),
                                      ((5, 3), 0.8), ((11, 11), 0), ((31, 31), 5.0)])
def test_gaussian_general_taps_u8(rcv, oracle, ks, sigma):
    R = rcv
    a = oracle.fill_u8(44, 61 * 83 * 3).reshape(61, 83, 3)
    d = R.Mat.empty()
    R.imgproc.gaussian_blur(R.Mat.from_numpy(a), d, ks, float(sigma))
    assert_same(d.to_numpy(), oracle.gaussian_blur(a, ks, float(sigma), float(sigma)), f"gauss {ks} {sigma}")


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("ks,sx,sy", [(3, 0.0, 0.0), (3, 0.6, 1.4), (5, 1.0, 1.0), (5, 0.7, 2.5), (7, 0.0, 0.0), (7, 1.5, 1.5),
                                      (7, 3.0, 0.9)])
def test_gaussian_q8_strip_kernel(rcv, oracle, cn, ks, sx, sy):
    """k_strip<GaussQ8Op<CN, KS>>: any-sigma 3/5/7 Gaussians, multi-strip, ragged right edge, band seams."""
    R = rcv
    h, w = 203, 517
    a = oracle.fill_u8(140 + cn + ks, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    s = mats(R, a, "device")
    d = s.like()
    R.imgproc.gaussian_blur(s, d, (ks, ks), sx, sy)
    assert_same(d.to_numpy(), oracle.gaussian_blur(a, (ks, ks), sx, sy), f"gaussq8 cn{cn} ks{ks} {sx} {sy}")


def test_gaussian_q8_extremes(rcv, oracle):
    """All-255 and checkerboard inputs sit on the 16-bit lane bounds of the packed arithmetic."""
    R = rcv
    for ks, sg in ((3, 0.0), (5, 1.3), (7, 2.0)):
        for a in (np.full((64, 200, 3), 255, np.uint8),
                  ((np.indices((64, 200)).sum(0) % 2) * 255).astype(np.uint8)[:, :, None].repeat(3, 2)):
            a = np.ascontiguousarray(a)
            s = mats(R, a, "device")
            d = s.like()
            R.imgproc.gaussian_blur(s, d, (ks, ks), sg)
            assert_same(d.to_numpy(), oracle.gaussian_blur(a, (ks, ks), sg, sg), f"extreme ks{ks}")
    a = np.full((64, 200, 3), 255, np.uint8)
    s = mats(R, a, "device")
    d = s.like()
    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
    assert (d.to_numpy() == 255).all()


def test_gaussian_strided_and_padding_untouched(rcv, oracle):
    """Mat.step > cols*channels on input AND output; padding bytes must survive."""
    R = rcv
    a = oracle.fill_u8(45, 70 * 90 * 3).reshape(70, 90, 3)
    s = R.Mat.from_numpy_strided(a, step=90 * 3 + 50)
    d = R.Mat.from_numpy_strided(np.zeros_like(a), step=90 * 3 + 13, fill=0x5A)
    # a strided host dst is used as it is when it already has the right geometry
    from rustcv_b200 import _ffi as F
    F.check(F.lib.rcv_gaussian_blur(C.byref(s.c()), C.byref(d.c()), 5, 5, 0.0, 0.0))
    assert_same(d.to_numpy(), oracle.gaussian_blur(a, (5, 5)), "strided")
    assert (d.data.reshape(70, d.step)[:, 270:] == 0x5A).all()


def test_gaussian_golden_fixtures(rcv, oracle, golden):
    """OpenCV 4.13 outputs committed under tests/golden/ (bit-exact)."""
    R = rcv
    H, W = [int(x) for x in golden["shape"]]
    bgr = oracle.fill_u8(7, H * W * 3).reshape(H, W, 3)
    for key in golden.files:
        if key.startswith("gauss_bgr_"):
            kw, kh, sg = key[len("gauss_bgr_"):].split("_")
            for where in ("host", "device"):
                s = mats(R, bgr, where)
                d = out_like(R, s, where)
                R.imgproc.gaussian_blur(s, d, (int(kw), int(kh)), float(sg))
                assert_same(d.to_numpy(), golden[key], f"{key} {where}")
    gray = oracle.fill_u8(8, H * W).reshape(H, W)
    bgra = oracle.fill_u8(10, H * W * 4).reshape(H, W, 4)
    for a, key in ((gray, "gauss_gray_5_5_0"), (bgra, "gauss_bgra_5_5_0")):
        s = mats(R, a, "device")
        d = s.like()
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        assert_same(d.to_numpy(), golden[key], key)


def test_gaussian_f32(rcv, oracle):
    R = rcv
    for cn in (1, 3):
        a = oracle.fill_f32(46, 61 * 83 * cn).reshape((61, 83) if cn == 1 else (61, 83, cn))
        for ks, sg in (((5, 5), 1.1), ((3, 3), 0.0), ((9, 7), 2.0)):
            d = R.Mat.empty()
            R.imgproc.gaussian_blur(R.Mat.from_numpy(a), d, ks, sg)
            want = oracle.gaussian_blur(a, ks, sg, sg)
            got = d.to_numpy()
            assert_f32(got.reshape(61, -1), want.reshape(61, -1), f"gauss f32 {ks} {sg} cn{cn}", max_ulp=0)


# ---- separable / dense filters -----------------------------------------------------------------
def test_sep_filter2d(rcv, oracle):
    R = rcv
    a = oracle.fill_f32(47, 50 * 77).reshape(50, 77)
    kx = np.array([0.25, 0.5, 0.25], np.float32)
    ky = np.array([-1, 0, 1, 2, 0.5], np.float32)
    d = R.Mat.empty()
    R.imgproc.sep_filter2d(R.Mat.from_numpy(a), d, kx, ky)
    assert_f32(d.to_numpy(), oracle.sepfilter_f32(a, kx, ky), "sepfilter f32", max_ulp=0)
    u = oracle.fill_u8(48, 50 * 77 * 3).reshape(50, 77, 3)
    qx = np.array([10, 50, 136, 50, 10], np.int32)
    qy = np.array([64, 128, 64], np.int32)
    R.imgproc.sep_filter2d(R.Mat.from_numpy(u), d, qx, qy)
    assert_same(d.to_numpy(), oracle.sepfilter_u8_q8(u, qx, qy), "sepfilter q8")


@pytest.mark.parametrize("ks", [3, 5, 7])
def test_f32_strip_kernels(rcv, oracle, ks):
    """k_strip<SepF32Op<KS>> / k_strip<Filter2dF32Op<KS>>: gray f32, multi-strip, ragged edge, seams."""
    R = rcv
    rng = np.random.default_rng(ks)
    h, w = 203, 517
    a = oracle.fill_f32(150 + ks, h * w).reshape(h, w)
    s = mats(R, a, "device")
    d = s.like()
    kx = rng.normal(size=ks).astype(np.float32)
    ky = rng.normal(size=ks).astype(np.float32)
    R.imgproc.sep_filter2d(s, d, kx, ky)
    assert_f32(d.to_numpy(), oracle.sepfilter_f32(a, kx, ky), f"sepf32 strip ks{ks}", max_ulp=0)
    R.imgproc.gaussian_blur(s, d, (ks, ks), 1.3)
    assert_f32(d.to_numpy(), oracle.gaussian_blur(a, (ks, ks), 1.3, 1.3), f"gauss f32 strip ks{ks}", max_ulp=0)
    if ks <= 5:
        k = rng.normal(size=(ks, ks)).astype(np.float32)
        R.imgproc.filter2d(s, d, k, delta=-0.5)
        assert_f32(d.to_numpy(), oracle.filter2d(a, k, -0.5), f"filter2d f32 strip ks{ks}", max_ulp=0)
    # the generic kernels agree
    R.imgproc.set_option("sepf32.force_generic", 1)
    R.imgproc.set_option("f2d.force_generic", 1)
    try:
        d2 = s.like()
        R.imgproc.sep_filter2d(s, d2, kx, ky)
        assert_f32(d2.to_numpy(), oracle.sepfilter_f32(a, kx, ky), f"sepf32 generic ks{ks}", max_ulp=0)
    finally:
        R.imgproc.set_option("sepf32.force_generic", 0)
        R.imgproc.set_option("f2d.force_generic", 0)


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
def test_filter2d_u8_3x3_strip_kernel(rcv, oracle, cn):
    """k_strip<Filter2dU8Op<CN,3>>: sharpen / Laplacian / random taps incl. exact .5 ties and saturation."""
    R = rcv
    rng = np.random.default_rng(cn)
    h, w = 203, 517
    a = oracle.fill_u8(160 + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    s = mats(R, a, "device")
    d = s.like()
    kernels = [np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32),           # sharpen (saturates both ways)
               np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32),               # Laplacian
               np.full((3, 3), 0.125, np.float32),                                     # exact binary fractions: .5 ties
               rng.normal(size=(3, 3)).astype(np.float32) / 3]
    for i, k in enumerate(kernels):
        for delta in (0.0, 0.5):
            R.imgproc.filter2d(s, d, k, delta=delta)
            assert_same(d.to_numpy(), oracle.filter2d(a, k, delta), f"filter2d u8 3x3 cn{cn} kernel{i} delta{delta}")


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", [(203, 517), (8, 8), (41, 163), (77, 1000)])
def test_filter2d_u8_5x5_strip_kernel(rcv, oracle, cn, shape):
    """k_strip<Filter2dU8Op<CN,5>> (transposed form: pending row sums instead of a window of rows): unsharp mask
    (saturates both ways), exact binary fractions (.5 ties round to even), an asymmetric ramp (catches any swap of tap
    order or direction), random taps; several bands and strips, the 8x8 minimum, ragged right edges; and the strip
    kernel against the general kernel."""
    R = rcv
    rng = np.random.default_rng(10 + cn)
    h, w = shape
    a = oracle.fill_u8(170 + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    s = mats(R, a, "device")
    d = s.like()
    unsharp = -np.ones((5, 5), np.float32) / 8
    unsharp[2, 2] = 4.0
    kernels = [unsharp, np.full((5, 5), 0.03125, np.float32), (np.arange(25, dtype=np.float32).reshape(5, 5) - 7) / 64,
               rng.normal(size=(5, 5)).astype(np.float32) / 5]
    n0 = R.imgproc.launch_count()
    for i, k in enumerate(kernels):
        for delta in (0.0, 0.5):
            R.imgproc.filter2d(s, d, k, delta=delta)
            assert_same(d.to_numpy(), oracle.filter2d(a, k, delta), f"filter2d u8 5x5 cn{cn} {shape} kernel{i} delta{delta}")
    assert R.imgproc.launch_count() - n0 == 8, "one strip launch per call"
    R.imgproc.set_option("f2d.force_generic", 1)
    try:
        d2 = s.like()
        R.imgproc.filter2d(s, d2, kernels[3], delta=0.5)
        assert_same(d2.to_numpy(), d.to_numpy(), "strip op vs general kernel")
    finally:
        R.imgproc.set_option("f2d.force_generic", 0)


@pytest.mark.parametrize("ksz", [(3, 3), (5, 5), (7, 3), (4, 4), (1, 1), (9, 9)])
def test_filter2d(rcv, oracle, ksz):
    R = rcv
    rng = np.random.default_rng(5)
    k = rng.normal(size=ksz).astype(np.float32)
    a = oracle.fill_f32(49, 45 * 67).reshape(45, 67)
    d = R.Mat.empty()
    R.imgproc.filter2d(R.Mat.from_numpy(a), d, k, delta=0.125)
    assert_f32(d.to_numpy(), oracle.filter2d(a, k, 0.125), f"filter2d f32 {ksz}", max_ulp=0)
    u = oracle.fill_u8(50, 45 * 67 * 3).reshape(45, 67, 3)
    ku = (k / max(1e-3, np.abs(k).sum())).astype(np.float32)
    R.imgproc.filter2d(R.Mat.from_numpy(u), d, ku, delta=3.0)
    assert_same(d.to_numpy(), oracle.filter2d(u, ku, 3.0), f"filter2d u8 {ksz}")


def test_filter2d_golden(rcv, oracle, golden):
    R = rcv
    H, W = [int(x) for x in golden["shape"]]
    bgr = oracle.fill_u8(7, H * W * 3).reshape(H, W, 3)
    lap = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    d = R.Mat.empty()
    R.imgproc.filter2d(R.Mat.from_numpy(bgr), d, lap / 3 + 0.2)
    assert_same(d.to_numpy(), golden["filter2d_bgr"], "filter2d cv2 golden")


# ---- Sobel + magnitude ----------------------------------------------------------------------------
@pytest.mark.parametrize("where", WHERE)
@pytest.mark.parametrize("shape", [(61, 83), (8, 8), (64, 120), (97, 121), (200, 500), (30, 1000), (1, 1), (1, 7), (7, 1),
                                   (5, 300), (300, 5)])
def test_sobel_magnitude(rcv, oracle, where, shape):
    R = rcv
    h, w = shape
    a = oracle.fill_f32(51 + h, h * w).reshape(h, w)
    want = oracle.sobel3(a, ("gx", "gy", "mag"))
    s = mats(R, a, where)
    mag, gx, gy = (out_like(R, s, where) for _ in range(3))
    R.imgproc.sobel_mag(s, mag, gx, gy)
    assert_f32(gx.to_numpy(), want["gx"], f"sobel gx {shape} {where}", max_ulp=0)
    assert_f32(gy.to_numpy(), want["gy"], f"sobel gy {shape} {where}", max_ulp=0)
    assert_f32(mag.to_numpy(), want["mag"], f"sobel mag {shape} {where}", max_ulp=1)
    # magnitude only (the fused config-3 form)
    m2 = out_like(R, s, where)
    R.imgproc.sobel_mag(s, m2)
    assert_f32(m2.to_numpy(), want["mag"], f"sobel mag-only {shape} {where}", max_ulp=1)


def test_sobel_cfg3_full_size(rcv, oracle, golden):
    """BASELINE.json config 3: 1920x1080 f32; plus the OpenCV tolerance pin."""
    R = rcv
    a = oracle.fill_f32(3, 1080 * 1920).reshape(1080, 1920)
    oracle.set_threads(8)
    want = oracle.sobel3(a, ("mag",))["mag"]
    oracle.set_threads(1)
    s = mats(R, a, "device")
    m = s.like()
    R.imgproc.sobel_mag(s, m)
    assert_f32(m.to_numpy(), want, "sobel cfg3", max_ulp=1)
    H, W = [int(x) for x in golden["shape"]]
    f = oracle.fill_f32(9, H * W).reshape(H, W)
    d = R.Mat.empty()
    R.imgproc.sobel_mag(R.Mat.from_numpy(f), d)
    assert np.abs(d.to_numpy() - golden["sobel_mag"]).max() < 4e-6


# ---- resize -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("case", [((61, 83, 3), (37, 64)), ((61, 83, 3), (122, 166)), ((61, 83, 3), (30, 41)),
                                  ((64, 96, 3), (16, 24)), ((128, 256, 3), (32, 64)), ((61, 83, 1), (15, 20)),
                                  ((40, 40, 4), (100, 7)), ((5, 5, 3), (1, 1)), ((1, 1, 3), (4, 6)), ((64, 64, 3), (32, 32))])
def test_resize_u8(rcv, oracle, where, case):
    R = rcv
    (h, w, cn), (dr, dc) = case
    a = oracle.fill_u8(60 + h, h * w * cn).reshape(h, w, cn)
    if cn == 1:
        a = a.reshape(h, w)
    s = mats(R, a, where)
    d = out_like(R, s, where, rows=dr, cols=dc)
    R.imgproc.resize(s, d, (dc, dr) if where == "host" else None)
    assert_same(d.to_numpy(), oracle.resize_bilinear(a, dr, dc), f"resize {case} {where}")


def test_resize_4x_fast_path_equals_general_model(rcv, oracle, golden):
    R = rcv
    big = oracle.fill_u8(11, 64 * 96 * 3).reshape(64, 96, 3)
    s = mats(R, big, "device")
    d = s.like(rows=16, cols=24)
    R.imgproc.resize(s, d)
    assert_same(d.to_numpy(), golden["resize4x_bgr"], "4x vs OpenCV golden")
    R.imgproc.set_option("resize.force_generic", 1)
    try:
        d2 = s.like(rows=16, cols=24)
        R.imgproc.resize(s, d2)
        assert_same(d2.to_numpy(), d.to_numpy(), "4x fast path vs general kernel")
    finally:
        R.imgproc.set_option("resize.force_generic", 0)


@pytest.mark.parametrize("case", [((300, 500, 3), (131, 277)), ((100, 150, 3), (333, 517)), ((256, 512, 3), (64, 128)),
                                  ((217, 1023, 1), (100, 400)), ((90, 2000, 3), (45, 21)), ((64, 333, 4), (64, 700)),
                                  ((33, 37, 3), (33, 37)), ((500, 81, 2), (17, 300))])
def test_resize_larger_geometries(rcv, oracle, case):
    """Several CTAs in both directions, up- and down-scaling, all channel counts, u8 and f32; the 4x fast path
    and the general kernel agree."""
    R = rcv
    (h, w, cn), (dr, dc) = case
    a = oracle.fill_u8(70 + h, h * w * cn).reshape(h, w, cn)
    if cn == 1:
        a = a.reshape(h, w)
    want = oracle.resize_bilinear(a, dr, dc)
    s = mats(R, a, "device")
    for opt in (None, "resize.force_generic"):
        if opt:
            R.imgproc.set_option(opt, 1)
        try:
            d = out_like(R, s, "device", rows=dr, cols=dc)
            R.imgproc.resize(s, d)
            assert_same(d.to_numpy(), want, f"resize {case} {opt}")
        finally:
            if opt:
                R.imgproc.set_option(opt, 0)
    f = oracle.fill_f32(71 + h, h * w * cn).reshape(a.shape)
    sf = mats(R, f, "device")
    df = out_like(R, sf, "device", rows=dr, cols=dc)
    R.imgproc.resize(sf, df)
    wantf = oracle.resize_bilinear(f, dr, dc)
    assert_f32(df.to_numpy().reshape(dr, -1), wantf.reshape(dr, -1), f"resize f32 {case}", max_ulp=0)


@pytest.mark.parametrize("cn", [1, 3, 4])
def test_resize_2x_fast_path(rcv, oracle, cn):
    """Exact 2x downscale (rounded 2x2 box mean): bit-identical to the oracle's fixed-point model, to the general
    kernel, and to OpenCV (cv2.resize INTER_LINEAR at exactly 2x; golden generated with the other fixtures)."""
    R = rcv
    for h, w in ((64, 96), (270, 480), (2, 16), (130, 1000)):
        a = oracle.fill_u8(80 + h + cn, h * w * cn).reshape(h, w, cn)
        if cn == 1:
            a = a.reshape(h, w)
        want = oracle.resize_bilinear(a, h // 2, w // 2)
        s = mats(R, a, "device")
        n0 = R.imgproc.launch_count()
        d = out_like(R, s, "device", rows=h // 2, cols=w // 2)
        n0 = R.imgproc.launch_count()
        R.imgproc.resize(s, d)
        assert R.imgproc.launch_count() - n0 == 1
        assert_same(d.to_numpy(), want, f"resize 2x cn{cn} {h}x{w}")
        box = (a.astype(np.uint32).reshape(h // 2, 2, w // 2, 2, -1).sum(axis=(1, 3)) + 2) >> 2
        assert_same(d.to_numpy().reshape(h // 2, w // 2, -1), box.astype(np.uint8), f"2x2 box mean cn{cn} {h}x{w}")
        R.imgproc.set_option("resize.force_generic", 1)
        try:
            d2 = out_like(R, s, "device", rows=h // 2, cols=w // 2)
            R.imgproc.resize(s, d2)
            assert_same(d2.to_numpy(), want, f"general kernel at 2x cn{cn} {h}x{w}")
        finally:
            R.imgproc.set_option("resize.force_generic", 0)
    # host Mats and a width that is not a multiple of the vector group (general kernel)
    a = oracle.fill_u8(90 + cn, 50 * 52 * cn).reshape(50, 52, cn)
    if cn == 1:
        a = a.reshape(50, 52)
    d = R.Mat.empty()
    R.imgproc.resize(R.Mat.from_numpy(a), d, (26, 25))
    assert_same(d.to_numpy(), oracle.resize_bilinear(a, 25, 26), f"resize 2x host cn{cn}")


@pytest.mark.parametrize("cn", [1, 3, 4])
def test_resize_2x_upscale_fast_path(rcv, oracle, cn):
    """Exact 2x upscale (k_resize_up2x_u8): the fixed-point model's closed form at scale 1/2 -- bit-identical to the
    oracle (itself bit-identical to OpenCV there), to the general kernel, and to the closed form written in numpy;
    edge rows / columns (clamped far taps), single-row and single-group images included."""
    R = rcv
    for h, w in ((32, 48), (135, 240), (1, 4), (7, 4), (1, 400), (61, 1000)):
        a = oracle.fill_u8(85 + h + cn, h * w * cn).reshape(h, w, cn)
        a3 = a.astype(np.int64)
        if cn == 1:
            a = a.reshape(h, w)
        want = oracle.resize_bilinear(a, 2 * h, 2 * w)
        dx, dy = np.arange(2 * w), np.arange(2 * h)
        far = np.where(dx % 2 == 0, np.maximum(dx // 2 - 1, 0), np.minimum(dx // 2 + 1, w - 1))
        farr = np.where(dy % 2 == 0, np.maximum(dy // 2 - 1, 0), np.minimum(dy // 2 + 1, h - 1))
        H = a3[:, far, :] + 3 * a3[:, dx // 2, :]
        closed = (((H[farr] >> 2) + ((3 * H[dy // 2]) >> 2) + 2) >> 2).astype(np.uint8)
        assert_same(closed.reshape(want.shape), want, f"closed form vs oracle cn{cn} {h}x{w}")
        s = mats(R, a, "device")
        d = out_like(R, s, "device", rows=2 * h, cols=2 * w)
        n0 = R.imgproc.launch_count()
        R.imgproc.resize(s, d)
        assert R.imgproc.launch_count() - n0 == 1
        assert_same(d.to_numpy(), want, f"resize up 2x cn{cn} {h}x{w}")
        R.imgproc.set_option("resize.force_generic", 1)
        try:
            d2 = out_like(R, s, "device", rows=2 * h, cols=2 * w)
            R.imgproc.resize(s, d2)
            assert_same(d2.to_numpy(), want, f"general kernel at 1/2 cn{cn} {h}x{w}")
        finally:
            R.imgproc.set_option("resize.force_generic", 0)
    a = oracle.fill_u8(95 + cn, 30 * 44 * cn).reshape(30, 44, cn)
    if cn == 1:
        a = a.reshape(30, 44)
    d = R.Mat.empty()
    R.imgproc.resize(R.Mat.from_numpy(a), d, (88, 60))
    assert_same(d.to_numpy(), oracle.resize_bilinear(a, 60, 88), f"resize up 2x host cn{cn}")
    # a batch (frames in blockIdx.z)
    frames = [oracle.fill_u8(96 + k, 40 * 64 * cn).reshape((40, 64, cn) if cn > 1 else (40, 64)) for k in range(3)]
    sb, db = R.Mat.device_batch(3, 40, 64, cn), R.Mat.device_batch(3, 80, 128, cn)
    for k in range(3):
        upload_into(R, frames[k], sb[k])
    R.imgproc.resize_batch(sb, db)
    for k in range(3):
        assert_same(db[k].to_numpy(), oracle.resize_bilinear(frames[k], 80, 128), f"resize up 2x batch {k} cn{cn}")
    sb.free(); db.free()


def test_resize_f32(rcv, oracle):
    R = rcv
    a = oracle.fill_f32(61, 61 * 83).reshape(61, 83)
    for dr, dc in ((30, 41), (100, 200), (61, 83)):
        d = R.Mat.empty()
        R.imgproc.resize(R.Mat.from_numpy(a), d, (dc, dr))
        assert_f32(d.to_numpy(), oracle.resize_bilinear(a, dr, dc), f"resize f32 {dr}x{dc}", max_ulp=0)


# ---- warpAffine -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("shape,angle,scale", [((61, 83), 15.0, 1.0), ((128, 128), 15.0, 1.0), ((50, 200), -37.0, 1.3),
                                               ((200, 50), 90.0, 1.0), ((64, 64), 0.0, 1.0), ((33, 47), 180.0, 0.5)])
def test_warp_affine_f32(rcv, oracle, where, shape, angle, scale):
    R = rcv
    h, w = shape
    a = oracle.fill_f32(70 + h, h * w).reshape(h, w)
    M = R.imgproc.get_rotation_matrix_2d(((w - 1) / 2, (h - 1) / 2), angle, scale)
    assert (M.ravel() == oracle.rotation_matrix((w - 1) / 2, (h - 1) / 2, angle, scale)).all()
    s = mats(R, a, where)
    d = out_like(R, s, where)
    R.imgproc.warp_affine(s, d, M, border_value=0.25)
    assert_f32(d.to_numpy(), oracle.warp_affine(a, M.ravel(), border_value=0.25), f"warp {shape} {angle} {where}", max_ulp=1)


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("angle,scale", [(15.0, 1.0), (-63.0, 0.8), (90.0, 1.0), (5.0, 2.5)])
def test_warp_affine_u8_tile_kernel(rcv, oracle, cn, angle, scale):
    """k_warp_tile<uint8_t, CN>: TMA-staged boxes in byte geometry, interior and border tiles."""
    R = rcv
    h, w = 150, 333
    a = oracle.fill_u8(190 + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    M = R.imgproc.get_rotation_matrix_2d(((w - 1) / 2, (h - 1) / 2), angle, scale)
    s = mats(R, a, "device")
    d = s.like()
    R.imgproc.warp_affine(s, d, M, border_value=9)
    assert_same(d.to_numpy(), oracle.warp_affine(a, M.ravel(), border_value=9), f"warp u8 cn{cn} {angle} {scale}")
    R.imgproc.set_option("warp.force_generic", 1)
    try:
        d2 = s.like()
        R.imgproc.warp_affine(s, d2, M, border_value=9)
        assert_same(d2.to_numpy(), d.to_numpy(), "tile kernel vs direct gather kernel")
    finally:
        R.imgproc.set_option("warp.force_generic", 0)


def test_warp_affine_u8_and_inverse_map(rcv, oracle):
    R = rcv
    a = oracle.fill_u8(71, 61 * 83 * 3).reshape(61, 83, 3)
    M = R.imgproc.get_rotation_matrix_2d((41.0, 30.0), 15.0, 1.1)
    d = R.Mat.empty()
    R.imgproc.warp_affine(R.Mat.from_numpy(a), d, M, dsize=(100, 70), border_value=7)
    assert_same(d.to_numpy(), oracle.warp_affine(a, M.ravel(), dsize=(70, 100), border_value=7), "warp u8")
    iM = oracle.invert_affine(M.ravel())
    R.imgproc.warp_affine(R.Mat.from_numpy(a), d, iM.reshape(2, 3), dsize=(100, 70), inverse_map=True, border_value=7)
    assert_same(d.to_numpy(), oracle.warp_affine(a, iM, dsize=(70, 100), inverse_map=True, border_value=7), "warp u8 inverse")


# ---- batches ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("where", ["device", "pinned", "host"])
def test_gaussian_batch(rcv, oracle, where):
    R = rcv
    n, h, w = 6, 131, 517
    frames = [oracle.fill_u8(80 + j, h * w * 3).reshape(h, w, 3) for j in range(n)]
    if where == "device":
        sb = R.Mat.device_batch(n, h, w, 3)
        db = R.Mat.device_batch(n, h, w, 3)
        from rustcv_b200 import _ffi as F
        for j in range(n):
            upload_into(R, frames[j], sb[j])
        before = R.imgproc.launch_count()
        R.imgproc.gaussian_blur_batch(sb, db)
        assert R.imgproc.launch_count() - before == 1, "a uniform device batch is ONE launch"
        outs = [db[j].to_numpy() for j in range(n)]
        sb.free()
        db.free()
    else:
        srcs = [mats(R, f, where) for f in frames]
        dsts = [s.like() for s in srcs]
        R.imgproc.gaussian_blur_batch(srcs, dsts)
        outs = [d.to_numpy() for d in dsts]
    for j in range(n):
        assert_same(outs[j], oracle.gaussian_blur(frames[j], (5, 5)), f"batch {where} frame {j}")


def test_other_batches(rcv, oracle):
    R = rcv
    n = 5
    # Sobel
    fr = [oracle.fill_f32(90 + j, 70 * 250).reshape(70, 250) for j in range(n)]
    sb = R.Mat.device_batch(n, 70, 250, 1, R.F32)
    db = R.Mat.device_batch(n, 70, 250, 1, R.F32)
    from rustcv_b200 import _ffi as F
    for j in range(n):
        upload_into(R, fr[j], sb[j])
    R.imgproc.sobel_mag_batch(sb, db)
    for j in range(n):
        assert_f32(db[j].to_numpy(), oracle.sobel3(fr[j])["mag"], f"sobel batch {j}", max_ulp=1)
    # warpAffine on the same frames
    M = R.imgproc.get_rotation_matrix_2d((124.5, 34.5), 15.0)
    R.imgproc.warp_affine_batch(sb, db, M)
    for j in range(n):
        assert_f32(db[j].to_numpy(), oracle.warp_affine(fr[j], M.ravel()), f"warp batch {j}", max_ulp=1)
    sb.free()
    db.free()
    # resize 4x, host batch through the staging ring (n > ring depth)
    fr = [oracle.fill_u8(95 + j, 64 * 128 * 3).reshape(64, 128, 3) for j in range(n)]
    srcs = [R.Mat.from_numpy(f) for f in fr]
    dsts = [R.Mat.new(16, 32, 3) for _ in fr]
    R.imgproc.resize_batch(srcs, dsts)
    for j in range(n):
        assert_same(dsts[j].to_numpy(), oracle.resize_bilinear(fr[j], 16, 32), f"resize batch {j}")
    # cvtColor batch
    yu = [oracle.fill_u8(99 + j, 48 * 64 * 2).reshape(48, 64, 2) for j in range(n)]
    srcs = [R.Mat.from_numpy(f) for f in yu]
    dsts = [R.Mat.new(48, 64, 3) for _ in yu]
    R.imgproc.cvt_color_batch(srcs, dsts, R.imgproc.COLOR_YUYV2BGR)
    for j in range(n):
        assert_same(dsts[j].to_numpy(), oracle.yuyv_to_bgr(yu[j]), f"cvt batch {j}")


def test_fused_yuyv_gaussian(rcv, oracle):
    R = rcv
    src = oracle.fill_u8(1, 480 * 640 * 2).reshape(480, 640, 2)
    d = R.Mat.empty()
    R.imgproc.yuyv_to_bgr_gaussian5(R.Mat.from_numpy(src), d)
    assert_same(d.to_numpy(), oracle.gaussian_blur(oracle.yuyv_to_bgr(src), (5, 5)), "yuyv->bgr->gauss5")


# ---- full BASELINE.json sizes: goldens and size-independent properties ---------------------------------------------
def test_cfg2_full_size_golden_crc(rcv, oracle):
    """3840x2160 BGR, SplitMix64 seed 2: dst CRC32 827081c8 (== OpenCV)."""
    R = rcv
    img = oracle.fill_u8(2, 2160 * 3840 * 3).reshape(2160, 3840, 3)
    for where in ("device", "host"):
        s = mats(R, img, where)
        d = out_like(R, s, where)
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
        got = d.to_numpy()
        assert oracle.crc32(got) == 0x827081C8, where
        assert got[0, 0].tolist() == [123, 104, 111] and got[1079, 1919].tolist() == [97, 153, 135]


def test_cfg2_properties_at_full_size(rcv, oracle):
    R = rcv
    h, w = 2160, 3840
    # constant image is a fixed point (taps sum to 256, single rounding)
    c = np.full((h, w, 3), 201, np.uint8)
    s = mats(R, c, "device")
    d = s.like()
    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
    assert (d.to_numpy() == 201).all()
    # channel independence + 180-degree symmetry: blur(flip(x)) == flip(blur(x))
    img = oracle.fill_u8(5, h * w * 3).reshape(h, w, 3)
    s = mats(R, img, "device")
    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
    a = d.to_numpy()
    s2 = mats(R, np.ascontiguousarray(img[::-1, ::-1]), "device")
    R.imgproc.gaussian_blur(s2, d, (5, 5), 0.0)
    assert (d.to_numpy()[::-1, ::-1] == a).all()
    s3 = mats(R, np.ascontiguousarray(img[:, :, ::-1]), "device")
    R.imgproc.gaussian_blur(s3, d, (5, 5), 0.0)
    assert (d.to_numpy()[:, :, ::-1] == a).all()
    # the fast kernel and the general kernel agree on the whole frame
    R.imgproc.set_option("gauss.force_generic", 1)
    try:
        R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
    finally:
        R.imgproc.set_option("gauss.force_generic", 0)
    assert (d.to_numpy() == a).all()


def test_cfg4_full_size_golden_crc(rcv, oracle):
    """7680x4320 -> 1920x1080 BGR, seed 4: dst CRC32 31a84a85 (== OpenCV INTER_LINEAR)."""
    R = rcv
    s = mats(R, oracle.fill_u8(4, 4320 * 7680 * 3).reshape(4320, 7680, 3), "device")
    d = s.like(rows=1080, cols=1920)
    R.imgproc.resize(s, d)
    assert oracle.crc32(d.to_numpy()) == 0x31A84A85


def test_cfg5_warp_full_size_properties(rcv, oracle):
    """4096x4096 f32, 15 degrees about the centre: rows sampled against the oracle."""
    R = rcv
    n = 4096
    a = oracle.fill_f32(5, n * n).reshape(n, n)
    assert oracle.crc32(a) == 0xC8A39EB9
    M = R.imgproc.get_rotation_matrix_2d(((n - 1) / 2, (n - 1) / 2), 15.0)
    s = mats(R, a, "device")
    d = s.like()
    R.imgproc.warp_affine(s, d, M)
    got = d.to_numpy()
    oracle.set_threads(8)
    want = oracle.warp_affine(a, M.ravel())
    oracle.set_threads(1)
    assert_f32(got, want, "warp cfg5", max_ulp=1)
    # identity warp is exact
    I = np.array([[1.0, 0, 0], [0, 1.0, 0]])
    R.imgproc.warp_affine(s, d, I)
    assert (d.to_numpy() == a).all()


# ---- API contract on the GPU -----------------------------------------------------------------------------
def test_mixed_locations_and_nonuniform_batches(rcv, oracle):
    """host src -> device dst, device src -> host dst, and device batches made of separate allocations."""
    R = rcv
    a = oracle.fill_u8(170, 90 * 200 * 3).reshape(90, 200, 3)
    want = oracle.gaussian_blur(a, (5, 5))
    h = R.Mat.from_numpy(a)
    d = R.Mat.device_new(90, 200, 3)
    R.imgproc.gaussian_blur(h, d, (5, 5), 0.0)
    assert_same(d.to_numpy(), want, "host -> device")
    back = R.Mat.empty()
    R.imgproc.gaussian_blur(h.upload(), back, (5, 5), 0.0)
    assert_same(back.to_numpy(), want, "device -> host")
    srcs, pads = [], []
    for j in range(3):  # three separate allocations with odd-sized allocations in between
        srcs.append(R.Mat.from_numpy(np.roll(a, j, axis=0)).upload())
        pads.append(R.Mat.device_new(7 + 13 * j, 100, 1))
    dsts = [s.like() for s in srcs]
    before = R.imgproc.launch_count()
    R.imgproc.gaussian_blur_batch(srcs, dsts)
    # one launch per frame unless the allocator happened to space the frames uniformly
    assert R.imgproc.launch_count() - before in (1, 3)
    for j in range(3):
        assert_same(dsts[j].to_numpy(), oracle.gaussian_blur(np.roll(a, j, axis=0), (5, 5)), f"non-uniform batch {j}")


def test_nonblocking_mode_and_sync(rcv, oracle):
    R = rcv
    a = oracle.fill_u8(171, 300 * 400 * 3).reshape(300, 400, 3)
    s = R.Mat.from_numpy(a).upload()
    d = s.like()
    R.imgproc.set_blocking(False)
    try:
        for _ in range(5):
            R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)  # enqueue only
        R.imgproc.sync()
    finally:
        R.imgproc.set_blocking(True)
    assert_same(d.to_numpy(), oracle.gaussian_blur(a, (5, 5)), "non-blocking")


def test_concurrent_callers(rcv, oracle):
    """The ABI is thread-safe for distinct Mats (the reference API is used from one thread per capture)."""
    import threading

    R = rcv
    imgs = [oracle.fill_u8(180 + t, 240 * 320 * 3).reshape(240, 320, 3) for t in range(4)]
    outs = [None] * 4
    errs = []

    def work(t):
        try:
            for _ in range(10):
                d = R.Mat.empty()
                R.imgproc.gaussian_blur(R.Mat.from_numpy(imgs[t]), d, (5, 5), 0.0)
                g = R.Mat.empty()
                R.imgproc.cvt_color(d, g, R.imgproc.COLOR_BGR2GRAY)
                outs[t] = (d.to_numpy(), g.to_numpy())
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for t in range(4):
        want = oracle.gaussian_blur(imgs[t], (5, 5))
        assert_same(outs[t][0], want, f"thread {t} blur")
        assert_same(outs[t][1], oracle.bgr_to_gray(want), f"thread {t} gray")


def test_errors_on_gpu(rcv):
    R = rcv
    from rustcv_b200 import _ffi as F

    s = R.Mat.device_new(16, 16, 3)
    small = R.Mat.device_new(8, 16, 3)
    with pytest.raises(F.RcvError) as e:
        R.imgproc.gaussian_blur(s, small, (5, 5), 0.0)  # device dst is never resized by the wrapper
    assert e.value.code == F.RCV_ERR_SIZE
    with pytest.raises(F.RcvError) as e:
        R.imgproc.gaussian_blur(s, s.like(), (4, 4), 0.0)  # even kernel size
    assert e.value.code == F.RCV_ERR_ARG
    with pytest.raises(F.RcvError) as e:
        R.imgproc.gaussian_blur(s, s.like(), (33, 33), 0.0)
    assert e.value.code == F.RCV_ERR_ARG
    fake = s.c()
    fake.device = 7  # not initialised
    assert F.lib.rcv_gaussian_blur(C.byref(fake), C.byref(s.like().c()), 5, 5, 0.0, 0.0) in (F.RCV_ERR_NOT_INIT, F.RCV_ERR_ARG)
    # the library is still healthy afterwards
    d = s.like()
    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)


# ---- randomized geometry sweep over the strip kernels --------------------------------------------------------
def test_strip_kernels_random_geometries(rcv, oracle):
    """60 random (rows, cols, channels, band height, location) draws per op family: every ragged-edge /
    band-seam / partial-chunk combination the strip pipeline can meet, against the oracle."""
    R = rcv
    rng = np.random.default_rng(20261017)
    for it in range(60):
        h = int(rng.integers(8, 260))
        w = int(rng.integers(8, 700))
        cn = int(rng.integers(1, 5))
        band = int(rng.choice([0, 4, 12, 20, 28, 36, 60]))
        where = str(rng.choice(["device", "host"]))
        a = rng.integers(0, 256, size=(h, w, cn), dtype=np.uint8)
        if cn == 1:
            a = a.reshape(h, w)
        R.imgproc.set_option("gauss.band_rows", band)
        try:
            s = mats(R, a, where)
            d = out_like(R, s, where)
            R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
            assert_same(d.to_numpy(), oracle.gaussian_blur(a, (5, 5)), f"rand gauss5 it{it} {h}x{w}x{cn} band{band} {where}")
            ks = int(rng.choice([3, 5, 7]))
            sg = float(rng.uniform(0.4, 3.0))
            d2 = out_like(R, s, where)
            R.imgproc.gaussian_blur(s, d2, (ks, ks), sg)
            assert_same(d2.to_numpy(), oracle.gaussian_blur(a, (ks, ks), sg, sg), f"rand gaussq8 it{it} {h}x{w}x{cn} ks{ks} s{sg:.2f} band{band}")
            k = rng.normal(size=(3, 3)).astype(np.float32) / 2
            R.imgproc.set_option("f2d.band_rows", band)
            d3 = out_like(R, s, where)
            R.imgproc.filter2d(s, d3, k, delta=1.5)
            assert_same(d3.to_numpy(), oracle.filter2d(a, k, 1.5), f"rand filter2d u8 it{it} {h}x{w}x{cn} band{band}")
        finally:
            R.imgproc.set_option("gauss.band_rows", 0)
            R.imgproc.set_option("f2d.band_rows", 0)
        # gray f32 family
        f = rng.random(size=(h, w), dtype=np.float32)
        R.imgproc.set_option("sobel.band_rows", band)
        R.imgproc.set_option("sepf32.band_rows", band)
        try:
            s = mats(R, f, where)
            m = out_like(R, s, where)
            R.imgproc.sobel_mag(s, m)
            assert_f32(m.to_numpy(), oracle.sobel3(f)["mag"], f"rand sobel it{it} {h}x{w} band{band} {where}", max_ulp=1)
            ks = int(rng.choice([3, 5, 7]))
            kx = rng.normal(size=ks).astype(np.float32)
            ky = rng.normal(size=ks).astype(np.float32)
            d = out_like(R, s, where)
            R.imgproc.sep_filter2d(s, d, kx, ky)
            assert_f32(d.to_numpy(), oracle.sepfilter_f32(f, kx, ky), f"rand sepf32 it{it} {h}x{w} ks{ks} band{band}", max_ulp=0)
        finally:
            R.imgproc.set_option("sobel.band_rows", 0)
            R.imgproc.set_option("sepf32.band_rows", 0)


def test_two_devices_in_one_process(rcv, oracle):
    """One context per GPU: Mats are routed by RcvMat.device (skipped on a single-GPU box)."""
    import ctypes as C2

    R = rcv
    from rustcv_b200 import _ffi as F

    n = C2.c_int()
    F.check(F.lib.rcv_device_count(C2.byref(n)))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    R.imgproc.init(1)
    a = oracle.fill_u8(200, 120 * 500 * 3).reshape(120, 500, 3)
    want = oracle.gaussian_blur(a, (5, 5))
    s1 = R.Mat.from_numpy(a).upload(device=1)
    assert s1.device == 1
    d1 = s1.like()
    R.imgproc.gaussian_blur(s1, d1, (5, 5), 0.0)
    assert_same(d1.to_numpy(), want, "device 1")
    s0 = R.Mat.from_numpy(a).upload(device=0)
    with pytest.raises(F.RcvError):
        R.imgproc.gaussian_blur(s0, d1, (5, 5), 0.0)  # Mats on different GPUs


def test_decode_frame_into_device_mat_then_process(rcv, oracle):
    """read() with a device-resident Mat: the raw YUYV frame is uploaded (2 B/px), converted on the GPU,
    and the BGR stays in HBM for the blur that follows; only the final result comes back."""
    R = rcv
    w, h = 640, 480
    raw = oracle.fill_u8(1, w * h * 2)
    frame = R.Mat.device_new(h, w, 3)
    assert R.videoio.decode_frame(raw, w, h, R.videoio.YUYV, frame)
    assert oracle.crc32(frame.to_numpy()) == 0x0BF66518
    blurred = frame.like()
    R.imgproc.gaussian_blur(frame, blurred, (5, 5), 0.0)
    want = oracle.gaussian_blur(oracle.yuyv_to_bgr(raw.reshape(h, w, 2)), (5, 5))
    assert_same(blurred.to_numpy(), want, "decode -> blur on device")
    # strided source rows (Frame.stride, ignored by the reference's facade)
    padded = np.full((h, w * 2 + 64), 0xEE, np.uint8)
    padded[:, : w * 2] = raw.reshape(h, w * 2)
    assert R.videoio.decode_frame(padded, w, h, R.videoio.YUYV, frame, stride=w * 2 + 64)
    assert oracle.crc32(frame.to_numpy()) == 0x0BF66518
    # BGR passthrough upload
    bgr = oracle.fill_u8(9, w * h * 3)
    assert R.videoio.decode_frame(bgr, w, h, R.videoio.BGR3, frame)
    assert (frame.to_numpy().ravel() == bgr).all()


@pytest.mark.parametrize("where", ["host", "device"])
def test_convert_to_and_gray_sobel_chain(rcv, oracle, where):
    """u8 <-> f32 conversion, and the capture-side chain BGR -> Gray -> f32 -> Sobel magnitude on one GPU."""
    R = rcv
    rng = np.random.default_rng(11)
    for shape in ((61, 83, 3), (40, 333), (5, 7, 4)):
        u = rng.integers(0, 256, size=shape, dtype=np.uint8)
        s = mats(R, u, where)
        d = out_like(R, s, where, depth=R.F32)
        R.imgproc.convert_to(s, d, R.F32, 1.0 / 255.0, 0.25)
        want = oracle.convert_to(u, np.float32, 1.0 / 255.0, 0.25)
        assert_f32(d.to_numpy().reshape(shape[0], -1), want.reshape(shape[0], -1), f"u8->f32 {shape}", max_ulp=0)
        f = (rng.random(size=shape, dtype=np.float32) * 300 - 20).astype(np.float32)
        s = mats(R, f, where)
        d = out_like(R, s, where, depth=R.U8)
        R.imgproc.convert_to(s, d, R.U8, 0.9, 3.5)
        assert_same(d.to_numpy(), oracle.convert_to(f, np.uint8, 0.9, 3.5), f"f32->u8 {shape}")
    bgr = oracle.fill_u8(210, 270 * 480 * 3).reshape(270, 480, 3)
    s = mats(R, bgr, where)
    g = out_like(R, s, where, channels=1)
    R.imgproc.cvt_color(s, g, R.imgproc.COLOR_BGR2GRAY)
    gf = out_like(R, g, where, depth=R.F32)
    R.imgproc.convert_to(g, gf, R.F32, 1.0 / 255.0)
    m = out_like(R, gf, where)
    R.imgproc.sobel_mag(gf, m)
    want = oracle.sobel3(oracle.convert_to(oracle.bgr_to_gray(bgr), np.float32, 1.0 / 255.0))["mag"]
    assert_f32(m.to_numpy(), want, f"gray->f32->sobel {where}", max_ulp=1)


# ---- fused decode -> process chain ---------------------------------------
def _yuyv_sobel_oracle(oracle, yuyv):
    """The chain's stand-alone stages on the CPU: reference BT.601 -> OpenCV gray -> f32 -> Sobel magnitude."""
    gray = oracle.bgr_to_gray(oracle.yuyv_to_bgr(yuyv))
    return oracle.sobel3(oracle.convert_to(gray, np.float32))["mag"]


@pytest.mark.parametrize("where", WHERE)
@pytest.mark.parametrize("shape", [(48, 240), (270, 480), (37, 254), (64, 722), (9, 8), (5, 6), (480, 640), (21, 1218)])
def test_yuyv_to_sobel_mag_fused(rcv, oracle, where, shape):
    """One kernel (YuyvSobelOp on the strip pipeline) against the four-stage oracle chain: exact integer stages
    and a correctly rounded sqrtf, so 0 ULP is demanded (the north star allows 1)."""
    R = rcv
    h, w = shape
    yuyv = oracle.fill_u8(300 + h + w, h * w * 2).reshape(h, w, 2)
    s = mats(R, yuyv, where)
    m = out_like(R, s, where, channels=1, depth=R.F32)
    R.imgproc.yuyv_to_sobel_mag(s, m)
    assert_f32(m.to_numpy(), _yuyv_sobel_oracle(oracle, yuyv), f"yuyv->sobel {shape} {where}", max_ulp=0)


def test_yuyv_to_sobel_mag_flat_and_saturated(rcv, oracle):
    """Flat regions give gx = gy = 0 (the one input outside the fast sqrt's range); saturated chroma exercises
    both clamp directions of the BT.601 stage."""
    R = rcv
    h, w = 40, 496
    yuyv = np.zeros((h, w, 2), np.uint8)
    yuyv[:, :, 0] = 128
    yuyv[:, :, 1] = 128
    yuyv[10:20, 100:300, 0] = 255
    yuyv[10:20, 100:300, 1] = 255
    yuyv[25:33, 8:40, :] = 0
    yuyv[:, 480:, 0] = np.arange(16, dtype=np.uint8) * 17
    s = R.Mat.from_numpy(yuyv).upload()
    m = s.like(channels=1, depth=R.F32)
    R.imgproc.yuyv_to_sobel_mag(s, m)
    got = m.to_numpy()
    assert_f32(got, _yuyv_sobel_oracle(oracle, yuyv), "yuyv->sobel flat", max_ulp=0)
    assert (got[0:8, 0:90] == 0).all()


def test_yuyv_to_sobel_mag_equals_unfused_library_chain(rcv, oracle):
    """The fused kernel and the library's own chain of stand-alone kernels (forced) agree bit for bit; padded
    host steps are honoured."""
    R = rcv
    h, w = 131, 1000
    yuyv = oracle.fill_u8(77, h * w * 2).reshape(h, w, 2)
    want = _yuyv_sobel_oracle(oracle, yuyv)
    s = R.Mat.from_numpy(yuyv).upload()
    a = s.like(channels=1, depth=R.F32)
    R.imgproc.yuyv_to_sobel_mag(s, a)
    n0 = R.imgproc.launch_count()
    R.imgproc.set_option("yuyvsobel.force_chain", 1)
    try:
        b = s.like(channels=1, depth=R.F32)
        R.imgproc.yuyv_to_sobel_mag(s, b)
    finally:
        R.imgproc.set_option("yuyvsobel.force_chain", 0)
    assert R.imgproc.launch_count() - n0 == 3, "the forced chain is three kernels (YUYV2GRAY, convertTo, Sobel)"
    n0 = R.imgproc.launch_count()
    c = s.like(channels=1, depth=R.F32)
    R.imgproc.yuyv_to_sobel_mag(s, c)
    assert R.imgproc.launch_count() - n0 == 1, "the fused path is ONE kernel"
    assert_f32(a.to_numpy(), want, "fused", max_ulp=0)
    assert_f32(b.to_numpy(), want, "chain", max_ulp=0)
    assert_f32(c.to_numpy(), want, "fused again", max_ulp=0)
    # host Mats with padded steps on both sides
    hs = R.Mat.from_numpy_strided(yuyv, w * 2 + 24)
    hd = R.Mat.from_numpy_strided(np.zeros((h, w), np.float32), w * 4 + 36)
    R.imgproc.yuyv_to_sobel_mag(hs, hd)
    assert_f32(hd.to_numpy(), want, "host padded", max_ulp=0)


def test_yuyv_to_sobel_mag_batch_and_errors(rcv, oracle):
    R = rcv
    n, h, w = 5, 72, 960
    src = R.Mat.device_batch(n, h, w, 2)
    dst = R.Mat.device_batch(n, h, w, 1, R.F32)
    frames = [oracle.fill_u8(500 + j, h * w * 2).reshape(h, w, 2) for j in range(n)]
    for j in range(n):
        upload_into(R, frames[j], src[j])
    n0 = R.imgproc.launch_count()
    R.imgproc.yuyv_to_sobel_mag_batch(src, dst)
    assert R.imgproc.launch_count() - n0 == 1
    for j in range(n):
        assert_f32(dst[j].to_numpy(), _yuyv_sobel_oracle(oracle, frames[j]), f"batch frame {j}", max_ulp=0)
    src.free(); dst.free()
    # host batch through the staging ring
    hs = [R.Mat.from_numpy(f) for f in frames]
    hd = [R.Mat.new(h, w, 1, R.F32) for _ in frames]
    R.imgproc.yuyv_to_sobel_mag_batch(hs, hd)
    for j in range(n):
        assert_f32(hd[j].to_numpy(), _yuyv_sobel_oracle(oracle, frames[j]), f"host batch frame {j}", max_ulp=0)
    # contract: odd widths, wrong channel counts and wrong destination geometry are errors, not silent returns
    from rustcv_b200 import _ffi as F
    odd = R.Mat.from_numpy(np.zeros((8, 9, 2), np.uint8))
    for bad_src, bad_dst, code in ((odd, R.Mat.new(8, 9, 1, R.F32), F.RCV_ERR_SIZE),
                                   (R.Mat.new(8, 8, 3), R.Mat.new(8, 8, 1, R.F32), F.RCV_ERR_DEPTH),
                                   (R.Mat.new(8, 8, 2), R.Mat.new(8, 8, 1), F.RCV_ERR_DEPTH),
                                   (R.Mat.new(8, 8, 2), R.Mat.new(8, 10, 1, R.F32), F.RCV_ERR_SIZE)):
        rc = F.lib.rcv_yuyv_to_sobel_mag(C.byref(bad_src.c()), C.byref(bad_dst.c()))
        assert rc == code, (rc, code, F.lib.rcv_last_error())


# ---- banded host pipeline: H2D / kernel / D2H of ONE frame overlap band by band ----------------------
@pytest.fixture
def small_bands(rcv):
    """Bands of 24 KB so that test-sized images are cut into many bands (default: 6 MB)."""
    rcv.imgproc.set_option("host.band_bytes", 24 << 10)
    yield rcv
    rcv.imgproc.set_option("host.band_bytes", 6 << 20)


def _pinned_pair(R, a, channels=None, depth=None):
    s = mats(R, a, "pinned")
    d = s.like(channels=channels, depth=depth)
    d.data[:] = 0x5A
    return s, d


def test_banded_pipeline_matches_oracle(small_bands, oracle):
    """Every op that declares a row window (strip kernels) or is pointwise, on pinned host Mats cut into
    8..40-row bands: identical to the oracle, band seams included."""
    R = small_bands
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, size=(333, 500, 3), dtype=np.uint8)
    for ks, sg in (((5, 5), 0.0), ((3, 3), 0.0), ((7, 7), 1.5), ((5, 5), 1.1)):
        s, d = _pinned_pair(R, bgr)
        n0 = R.imgproc.launch_count()
        R.imgproc.gaussian_blur(s, d, ks, sg)
        assert R.imgproc.launch_count() - n0 > 4, "expected one launch per band"
        assert_same(d.to_numpy(), oracle.gaussian_blur(bgr, ks, sg), f"banded gaussian {ks} {sg}")
    gray = rng.integers(0, 256, size=(301, 777), dtype=np.uint8)
    s, d = _pinned_pair(R, gray)
    R.imgproc.gaussian_blur(s, d, (5, 5), 0.0)
    assert_same(d.to_numpy(), oracle.gaussian_blur(gray, (5, 5), 0.0), "banded gaussian gray")
    f = rng.random(size=(257, 640), dtype=np.float32)
    s, d = _pinned_pair(R, f)
    R.imgproc.sobel_mag(s, d)
    assert_f32(d.to_numpy(), oracle.sobel3(f)["mag"], "banded sobel", max_ulp=1)
    kx = np.array([0.25, 0.5, 0.25], np.float32)
    ky = np.array([0.1, 0.2, 0.4, 0.2, 0.1], np.float32)
    for k1, k2 in ((kx, kx), (ky, ky)):
        s, d = _pinned_pair(R, f)
        R.imgproc.sep_filter2d(s, d, k1, k2)
        assert_f32(d.to_numpy(), oracle.sepfilter_f32(f, k1, k2), "banded sepfilter f32", max_ulp=0)
    k33 = rng.random(size=(3, 3), dtype=np.float32) - 0.4
    s, d = _pinned_pair(R, f)
    R.imgproc.filter2d(s, d, k33, 0.25)
    assert_f32(d.to_numpy(), oracle.filter2d(f, k33, 0.25), "banded filter2d f32", max_ulp=0)
    s, d = _pinned_pair(R, bgr)
    R.imgproc.filter2d(s, d, k33, 1.0)
    assert_same(d.to_numpy(), oracle.filter2d(bgr, k33, 1.0), "banded filter2d u8")
    # pointwise ops: bands are sub-images
    yuyv = rng.integers(0, 256, size=(240, 322, 2), dtype=np.uint8)
    s, d = _pinned_pair(R, yuyv, channels=3)
    R.imgproc.cvt_color(s, d, R.imgproc.COLOR_YUYV2BGR)
    assert_same(d.to_numpy(), oracle.yuyv_to_bgr(yuyv), "banded yuyv->bgr")
    s, d = _pinned_pair(R, bgr, channels=1)
    R.imgproc.cvt_color(s, d, R.imgproc.COLOR_BGR2GRAY)
    assert_same(d.to_numpy(), oracle.bgr_to_gray(bgr), "banded bgr->gray")
    s, d = _pinned_pair(R, bgr, depth=R.F32)
    R.imgproc.convert_to(s, d, R.F32, 0.5, 1.0)
    assert_f32(d.to_numpy().reshape(333, -1), oracle.convert_to(bgr, np.float32, 0.5, 1.0).reshape(333, -1), "banded convertTo", max_ulp=0)
    s, d = _pinned_pair(R, yuyv, channels=1, depth=R.F32)
    R.imgproc.yuyv_to_sobel_mag(s, d)
    assert_f32(d.to_numpy(), _yuyv_sobel_oracle(oracle, yuyv), "banded yuyv->sobel", max_ulp=0)


def test_banded_pipeline_fallback_and_mixed_locations(small_bands, oracle):
    """A kernel size only the whole-image kernels cover (9x9) silently takes the unbanded path; pinned -> device
    and device -> pinned pairs band one side only; batches of pinned frames pipeline frame by frame (unbanded)."""
    R = small_bands
    rng = np.random.default_rng(6)
    bgr = rng.integers(0, 256, size=(200, 400, 3), dtype=np.uint8)
    s, d = _pinned_pair(R, bgr)
    R.imgproc.gaussian_blur(s, d, (9, 9), 2.0)
    assert_same(d.to_numpy(), oracle.gaussian_blur(bgr, (9, 9), 2.0), "9x9 falls back to the unbanded path")
    want = oracle.gaussian_blur(bgr, (5, 5), 0.0)
    dev_dst = R.Mat.from_numpy(np.zeros_like(bgr)).upload()
    R.imgproc.gaussian_blur(s, dev_dst, (5, 5), 0.0)
    assert_same(dev_dst.to_numpy(), want, "pinned -> device")
    dev_src = R.Mat.from_numpy(bgr).upload()
    _, d2 = _pinned_pair(R, bgr)
    R.imgproc.gaussian_blur(dev_src, d2, (5, 5), 0.0)
    assert_same(d2.to_numpy(), want, "device -> pinned")
    frames = [rng.integers(0, 256, size=(150, 320, 3), dtype=np.uint8) for _ in range(7)]
    hs = [mats(R, f, "pinned") for f in frames]
    hd = [m.like() for m in hs]
    R.imgproc.gaussian_blur_batch(hs, hd, (5, 5), 0.0)
    for j, f in enumerate(frames):
        assert_same(hd[j].to_numpy(), oracle.gaussian_blur(f, (5, 5), 0.0), f"pinned batch frame {j}")
    # direct write: the kernel stores straight into the pinned destination (optional path, off by default)
    R.imgproc.set_option("host.direct_write", 1)
    try:
        _, d3 = _pinned_pair(R, bgr)
        R.imgproc.gaussian_blur(s, d3, (5, 5), 0.0)
        assert_same(d3.to_numpy(), want, "pinned, direct write")
        R.imgproc.set_option("host.zero_copy", 1)
        _, d4 = _pinned_pair(R, bgr)
        R.imgproc.gaussian_blur(s, d4, (5, 5), 0.0)
        assert_same(d4.to_numpy(), want, "pinned, zero copy both ways")
    finally:
        R.imgproc.set_option("host.direct_write", 0)
        R.imgproc.set_option("host.zero_copy", 0)
    # pageable Mats never band (the driver stages those copies itself)
    n0 = R.imgproc.launch_count()
    hp = R.Mat.empty()
    R.imgproc.gaussian_blur(R.Mat.from_numpy(bgr), hp, (5, 5), 0.0)
    assert R.imgproc.launch_count() - n0 == 1
    assert_same(hp.to_numpy(), want, "pageable")


# ---- fused YUYV -> BGR -> GaussianBlur 5x5 (one kernel) ------------------------------------------------
def _yuyv_gauss_oracle(oracle, yuyv):
    return oracle.gaussian_blur(oracle.yuyv_to_bgr(yuyv), (5, 5))


# widths chosen so the row ends at every macro-pixel phase of a lane (m = 0..3), exactly on a lane boundary,
# exactly on a strip boundary (240 px), one macro-pixel into the next strip, and inside the left-edge lane
YG_SHAPES = [(40, 8), (40, 10), (40, 12), (40, 14), (40, 16), (33, 238), (33, 240), (33, 242), (33, 244), (33, 246),
             (33, 248), (9, 480), (17, 482), (270, 480), (64, 1000), (480, 640), (11, 722)]


@pytest.mark.parametrize("where", ["device", "host"])
@pytest.mark.parametrize("shape", YG_SHAPES)
def test_yuyv_to_bgr_gaussian5_fused(rcv, oracle, where, shape):
    R = rcv
    h, w = shape
    yuyv = oracle.fill_u8(700 + h + w, h * w * 2).reshape(h, w, 2)
    s = mats(R, yuyv, where)
    d = out_like(R, s, where, channels=3)
    n0 = R.imgproc.launch_count()
    R.imgproc.yuyv_to_bgr_gaussian5(s, d)
    assert R.imgproc.launch_count() - n0 == 1, "one fused kernel"
    assert_same(d.to_numpy(), _yuyv_gauss_oracle(oracle, yuyv), f"yuyv->bgr->gauss5 fused {shape} {where}")


def test_yuyv_to_bgr_gaussian5_chain_batch_bands(rcv, oracle):
    """The fused kernel equals the library's two stand-alone kernels (forced), on batches, across internal band
    seams, with saturated inputs, and on tiny / odd-width images that take the two-kernel path by themselves."""
    R = rcv
    h, w = 211, 1202
    yuyv = oracle.fill_u8(78, h * w * 2).reshape(h, w, 2)
    yuyv[50:90, 300:700] = 255
    yuyv[100:140, 0:64] = 0
    want = _yuyv_gauss_oracle(oracle, yuyv)
    s = R.Mat.from_numpy(yuyv).upload()
    for br in (0, 8, 12, 36, 100):
        R.imgproc.set_option("yuyvgauss.band_rows", br)
        d = s.like(channels=3)
        R.imgproc.yuyv_to_bgr_gaussian5(s, d)
        assert_same(d.to_numpy(), want, f"fused, band_rows {br}")
    R.imgproc.set_option("yuyvgauss.band_rows", 0)
    R.imgproc.set_option("yuyvgauss.force_chain", 1)
    try:
        d = s.like(channels=3)
        n0 = R.imgproc.launch_count()
        R.imgproc.yuyv_to_bgr_gaussian5(s, d)
        assert R.imgproc.launch_count() - n0 == 2
        assert_same(d.to_numpy(), want, "forced two-kernel chain")
    finally:
        R.imgproc.set_option("yuyvgauss.force_chain", 0)
    for shape in ((5, 6), (7, 30), (12, 34)):
        y2 = oracle.fill_u8(79, shape[0] * shape[1] * 2).reshape(shape[0], shape[1], 2)
        d = R.Mat.empty()
        R.imgproc.yuyv_to_bgr_gaussian5(R.Mat.from_numpy(y2), d)
        assert_same(d.to_numpy(), _yuyv_gauss_oracle(oracle, y2), f"small {shape}")
    # an odd width would leave the BGR intermediate's last column unconverted (videoio/mod.rs:350) for the blur to
    # read: rejected, like rcv_yuyv_to_sobel_mag
    from rustcv_b200 import _ffi as F
    y3 = oracle.fill_u8(79, 12 * 33 * 2).reshape(12, 33, 2)
    with pytest.raises(F.RcvError) as e:
        R.imgproc.yuyv_to_bgr_gaussian5(R.Mat.from_numpy(y3), R.Mat.empty())
    assert e.value.code == F.RCV_ERR_SIZE
    n, hh, ww = 4, 96, 720
    frames = [oracle.fill_u8(800 + j, hh * ww * 2).reshape(hh, ww, 2) for j in range(n)]
    src = R.Mat.device_batch(n, hh, ww, 2)
    dst = R.Mat.device_batch(n, hh, ww, 3)
    for j in range(n):
        upload_into(R, frames[j], src[j])
    n0 = R.imgproc.launch_count()
    R.imgproc.yuyv_to_bgr_gaussian5_batch(src, dst)
    assert R.imgproc.launch_count() - n0 == 1
    for j in range(n):
        assert_same(dst[j].to_numpy(), _yuyv_gauss_oracle(oracle, frames[j]), f"batch frame {j}")
    src.free(); dst.free()
    hs = [mats(R, f, "pinned") for f in frames]
    hd = [m.like(channels=3) for m in hs]
    R.imgproc.yuyv_to_bgr_gaussian5_batch(hs, hd)
    for j in range(n):
        assert_same(hd[j].to_numpy(), _yuyv_gauss_oracle(oracle, frames[j]), f"pinned batch frame {j}")


# ---- MJPEG branch of read() through nvJPEG ------------------------------------------------------------
def test_mjpeg_decode_matches_libjpeg_turbo_within_tolerance(rcv):
    """rcv_mjpeg_to_bgr (nvJPEG entropy decode + IDCT, then this repo's libjpeg-style fancy upsampling and colour
    conversion kernel) against cv2.imdecode (libjpeg-turbo, the reference's decoder family) on the committed JPEG
    frames: 4:4:4, 4:2:0, 4:2:2, odd sizes, grayscale.  Decoders differ in the IDCT's last bit, so this is a
    tolerance, not bit-exactness: every sample within 4 levels, mean absolute difference below 0.1 (measured:
    max 3, mean 0.013-0.035: about 3 % of the samples are off by one).  With nvJPEG's own colour path (option
    mjpeg.library_color) the same frames differ by a mean of 0.5 (4:4:4) to 2.2 levels and up to 88 levels at
    chroma edges (it replicates chroma; TurboJPEG's default is the triangle filter) -- the reason the upsampling
    and the conversion are done by this repo's kernel."""
    import os

    R = rcv
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mjpeg_golden.npz"))
    stats, lib_stats = {}, {}
    names = ("444_q95", "420_q90", "422_q85", "420_odd_q92", "422_odd_q92", "gray_q90")
    for name in names:
        jpeg, want = g[f"jpeg_{name}"], g[f"bgr_{name}"]
        h, w = want.shape[:2]
        assert R.videoio.mjpeg_info(jpeg) == (w, h)
        host = R.Mat.empty()
        assert R.videoio.decode_frame(jpeg, 0, 0, R.videoio.MJPEG, host)
        assert (host.rows, host.cols, host.channels, host.step) == (h, w, 3, w * 3)
        dev = R.Mat.device_new(h, w, 3)
        assert R.videoio.decode_frame(jpeg, w, h, R.videoio.MJPEG, dev)
        got = host.to_numpy()
        assert (dev.to_numpy() == got).all(), "host and device destinations hold the same decode"
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        stats[name] = (int(d.max()), round(float(d.mean()), 3))
        R.imgproc.set_option("mjpeg.library_color", 1)
        try:
            lib = R.Mat.empty()
            R.videoio.decode_frame(jpeg, 0, 0, R.videoio.MJPEG, lib)
        finally:
            R.imgproc.set_option("mjpeg.library_color", 0)
        dl = np.abs(lib.to_numpy().astype(np.int32) - want.astype(np.int32))
        lib_stats[name] = (int(dl.max()), round(float(dl.mean()), 3))
    print("mjpeg vs libjpeg-turbo (max, mean):", stats)
    print("nvJPEG's own colour path   (max, mean):", lib_stats)
    for name, (mx, mean) in stats.items():
        assert mx <= 4, (name, mx)
        assert mean <= 0.1, (name, mean)
    for name, (mx, mean) in lib_stats.items():
        assert mean <= 4.0, (name, mean)
    # contract: garbage and wrongly sized destinations are errors
    from rustcv_b200 import _ffi as F
    junk = np.arange(64, dtype=np.uint8)
    m = R.Mat.new(8, 8, 3)
    assert F.lib.rcv_mjpeg_to_bgr(junk.ctypes.data, junk.size, C.byref(m.c())) == F.RCV_ERR_ARG
    jpeg = g["jpeg_444_q95"]
    assert F.lib.rcv_mjpeg_to_bgr(jpeg.ctypes.data, jpeg.size, C.byref(m.c())) == F.RCV_ERR_SIZE
    g1 = R.Mat.new(64, 96, 1)
    assert F.lib.rcv_mjpeg_to_bgr(jpeg.ctypes.data, jpeg.size, C.byref(g1.c())) == F.RCV_ERR_DEPTH


def test_fused_chains_random_geometries(rcv, oracle):
    """40 random (rows, even cols, band height, location) draws for both fused decode -> process kernels: every
    right-edge phase, band seam and partial chunk they can meet, against the oracle's stand-alone stages."""
    R = rcv
    rng = np.random.default_rng(20261018)
    for it in range(40):
        h = int(rng.integers(8, 200))
        w = 2 * int(rng.integers(4, 420))
        band = int(rng.choice([0, 4, 12, 20, 36, 60]))
        where = str(rng.choice(["device", "host", "pinned"]))
        y = rng.integers(0, 256, size=(h, w, 2), dtype=np.uint8)
        R.imgproc.set_option("yuyvgauss.band_rows", band)
        R.imgproc.set_option("yuyvsobel.band_rows", band)
        try:
            s = mats(R, y, where)
            d = out_like(R, s, where, channels=3)
            R.imgproc.yuyv_to_bgr_gaussian5(s, d)
            assert_same(d.to_numpy(), _yuyv_gauss_oracle(oracle, y), f"rand yuyv-gauss it{it} {h}x{w} band{band} {where}")
            m = out_like(R, s, where, channels=1, depth=R.F32)
            R.imgproc.yuyv_to_sobel_mag(s, m)
            assert_f32(m.to_numpy(), _yuyv_sobel_oracle(oracle, y), f"rand yuyv-sobel it{it} {h}x{w} band{band} {where}", max_ulp=0)
        finally:
            R.imgproc.set_option("yuyvgauss.band_rows", 0)
            R.imgproc.set_option("yuyvsobel.band_rows", 0)


def test_fused_chains_full_size_equal_the_unfused_kernels(rcv, oracle):
    """At camera sizes (4K and 1080p YUYV frames) the oracle chain takes a while, so the size-independent property
    is used: the fused kernel and the library's stand-alone kernels run back to back give identical bytes; one
    1080p frame is also checked against the oracle."""
    R = rcv
    for h, w in ((2160, 3840), (1080, 1920)):
        y = oracle.fill_u8(900 + h, h * w * 2).reshape(h, w, 2)
        s = R.Mat.from_numpy(y).upload()
        a, b = s.like(channels=3), s.like(channels=3)
        R.imgproc.yuyv_to_bgr_gaussian5(s, a)
        R.imgproc.set_option("yuyvgauss.force_chain", 1)
        try:
            R.imgproc.yuyv_to_bgr_gaussian5(s, b)
        finally:
            R.imgproc.set_option("yuyvgauss.force_chain", 0)
        an = a.to_numpy()
        assert_same(an, b.to_numpy(), f"fused vs unfused gauss chain {h}x{w}")
        ma, mb = s.like(channels=1, depth=R.F32), s.like(channels=1, depth=R.F32)
        R.imgproc.yuyv_to_sobel_mag(s, ma)
        R.imgproc.set_option("yuyvsobel.force_chain", 1)
        try:
            R.imgproc.yuyv_to_sobel_mag(s, mb)
        finally:
            R.imgproc.set_option("yuyvsobel.force_chain", 0)
        assert_f32(ma.to_numpy(), mb.to_numpy(), f"fused vs unfused sobel chain {h}x{w}", max_ulp=0)
        if h == 1080:
            oracle.set_threads(8)
            try:
                assert_same(an, _yuyv_gauss_oracle(oracle, y), "1080p fused gauss chain vs oracle")
                assert_f32(ma.to_numpy(), _yuyv_sobel_oracle(oracle, y), "1080p fused sobel chain vs oracle", max_ulp=0)
            finally:
                oracle.set_threads(1)
