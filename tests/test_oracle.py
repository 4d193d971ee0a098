"""CPU suite, part 1: pins the oracle.

(a) the reference's own unit tests for the pixel path (rustcv-camera/src/decode.rs:234-273),
(b) hand-derived known answers and SplitMix64 frame CRCs from SURVEY.md section 8c/8d,
(c) OpenCV 4.13 outputs (tests/golden/cv2_golden.npz) for the ops the reference lacks,
(d) an independent numpy restatement of the BT.601 formula over the whole (Y,U,V) domain.
"""
import numpy as np
import pytest


# ---- (a) reference unit tests --------------------------------------------------------
def test_ref_yuyv_to_bgr_basic(oracle):
    # decode.rs:235-252: Y=235, U=V=128 -> every channel > 240 (exactly 255)
    src = np.array([235, 128, 235, 128], np.uint8)
    for fn in (oracle.yuyv_to_bgr_facade, oracle.yuyv_to_bgr_camera):
        st, dst = fn(src, 2, 1)
        assert st == 0 and (dst > 240).all() and (dst == 255).all()


def test_ref_yuyv_to_bgr_black(oracle):
    # decode.rs:255-265: Y=16 -> every channel < 10 (exactly 0)
    src = np.array([16, 128, 16, 128], np.uint8)
    st, dst = oracle.yuyv_to_bgr_camera(src, 2, 1)
    assert st == 0 and (dst < 10).all() and (dst == 0).all()


def test_ref_rgb_to_bgr_swap(oracle):
    # decode.rs:268-273
    src = np.array([255, 0, 0, 0, 255, 0], np.uint8).reshape(1, 2, 3)
    assert oracle.swap_rb(src).ravel().tolist() == [0, 0, 255, 0, 255, 0]


# ---- (b) known answers / goldens ---------------------------------------------------------
def test_splitmix64_check_value(oracle):
    assert int(oracle.splitmix64(0, 1)[0]) == 0xE220A8397B1DCDAF
    # C and numpy generators agree
    import ctypes as C
    buf = np.zeros(1001, np.uint8)
    oracle.lib().orc_fill_u8(C.c_uint64(5), C.c_void_p(buf.ctypes.data), C.c_size_t(buf.size))
    assert (buf == oracle.fill_u8(5, 1001)).all()
    fb = np.zeros(77, np.float32)
    oracle.lib().orc_fill_f32(C.c_uint64(6), C.c_void_p(fb.ctypes.data), C.c_size_t(fb.size))
    assert (fb == oracle.fill_f32(6, 77)).all()


def test_yuyv_known_answers(oracle):
    st, d = oracle.yuyv_to_bgr_facade(np.full(4, 255, np.uint8), 2, 1)
    assert d.tolist() == [255, 125, 255, 255, 125, 255]
    st, d = oracle.yuyv_to_bgr_facade(np.zeros(4, np.uint8), 2, 1)
    assert d.tolist() == [0, 135, 0, 0, 135, 0]


def test_yuyv_cfg1_golden(oracle):
    src = oracle.fill_u8(1, 640 * 480 * 2)
    assert oracle.crc32(src) == 0x08EB63F6
    st, dst = oracle.yuyv_to_bgr_facade(src, 640, 480)
    assert st == 0 and oracle.crc32(dst) == 0x0BF66518
    assert dst[:6].tolist() == [133, 213, 220, 0, 0, 0]
    # strided form == packed form on an even-width frame
    assert (oracle.yuyv_to_bgr(src.reshape(480, 640, 2)).ravel() == dst).all()


def test_yuyv_contract_edges(oracle):
    src = oracle.fill_u8(3, 6 * 4 * 2)
    # facade: short src -> silent return (1); short dst -> would panic (-1)
    assert oracle.yuyv_to_bgr_facade(src[:-1], 6, 4)[0] == 1
    assert oracle.yuyv_to_bgr_facade(src, 6, 4, dst_len=6 * 4 * 3 - 1)[0] == -1
    # camera crate checks both
    assert oracle.yuyv_to_bgr_camera(src, 6, 4, dst_len=6 * 4 * 3 - 1)[0] == 1
    # odd pixel count: last pixel dropped (w*h/2 iterations)
    st, d = oracle.yuyv_to_bgr_facade(oracle.fill_u8(3, 3 * 3 * 2), 3, 3)
    assert st == 0 and (d[24:] == 0).all()


def test_yuv_formula_exhaustive(oracle):
    """All 2^24 (Y,U,V): C oracle == an independent numpy statement of videoio/mod.rs:352-369."""
    u, v = np.meshgrid(np.arange(256, dtype=np.int32), np.arange(256, dtype=np.int32), indexing="ij")
    u = u.ravel()
    v = v.ravel()
    for y in range(0, 256):
        frame = np.empty((65536, 4), np.uint8)
        frame[:, 0] = y
        frame[:, 1] = u
        frame[:, 2] = 255 - y
        frame[:, 3] = v
        st, got = oracle.yuyv_to_bgr_facade(frame.ravel(), 65536 * 2, 1)
        got = got.reshape(65536, 6)
        for k, yy in ((0, y), (3, 255 - y)):
            c = yy - 16
            b = np.clip((298 * c + 516 * (u - 128) + 128) >> 8, 0, 255)
            g = np.clip((298 * c - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8, 0, 255)
            r = np.clip((298 * c + 409 * (v - 128) + 128) >> 8, 0, 255)
            assert (got[:, k] == b).all() and (got[:, k + 1] == g).all() and (got[:, k + 2] == r).all()


def test_gauss_cfg2_golden(oracle):
    oracle.set_threads(8)
    img = oracle.fill_u8(2, 2160 * 3840 * 3).reshape(2160, 3840, 3)
    assert oracle.crc32(img) == 0xD15B0894
    g = oracle.gaussian_blur(img, (5, 5))
    oracle.set_threads(1)
    assert oracle.crc32(g) == 0x827081C8
    assert g[0, 0].tolist() == [123, 104, 111] and g[1079, 1919].tolist() == [97, 153, 135]


def test_sobel_cfg3_input_golden(oracle):
    f = oracle.fill_f32(3, 1080 * 1920)
    assert oracle.crc32(f) == 0xA837E897
    assert np.allclose(f[:3], [0.85549253, 0.11345029, 0.4824472], rtol=0, atol=1e-8)


def test_resize_cfg4_golden(oracle):
    oracle.set_threads(8)
    s = oracle.fill_u8(4, 4320 * 7680 * 3).reshape(4320, 7680, 3)
    assert oracle.crc32(s) == 0xC7614ADC
    r = oracle.resize_bilinear(s, 1080, 1920)
    oracle.set_threads(1)
    assert oracle.crc32(r) == 0x31A84A85
    # exact-4x closed form
    a = s[1::4, 1::4].astype(np.int32) + s[1::4, 2::4] + s[2::4, 1::4] + s[2::4, 2::4]
    assert (((a + 2) >> 2).astype(np.uint8) == r).all()


def test_threads_do_not_change_results(oracle):
    img = oracle.fill_u8(12, 97 * 131 * 3).reshape(97, 131, 3)
    a = oracle.gaussian_blur(img, (7, 7), 1.3)
    oracle.set_threads(5)
    b = oracle.gaussian_blur(img, (7, 7), 1.3)
    oracle.set_threads(1)
    assert (a == b).all()


# ---- (c) OpenCV pins -------------------------------------------------------------------------
def _inputs(oracle, golden):
    H, W = [int(x) for x in golden["shape"]]
    return (oracle.fill_u8(7, H * W * 3).reshape(H, W, 3), oracle.fill_u8(8, H * W).reshape(H, W),
            oracle.fill_f32(9, H * W).reshape(H, W), oracle.fill_u8(10, H * W * 4).reshape(H, W, 4))


def test_cv2_gaussian_bit_exact(oracle, golden):
    bgr, gray, _, bgra = _inputs(oracle, golden)
    for key in golden.files:
        if not key.startswith("gauss_bgr_"):
            continue
        kw, kh, sg = key[len("gauss_bgr_"):].split("_")
        got = oracle.gaussian_blur(bgr, (int(kw), int(kh)), float(sg), float(sg))
        assert (got == golden[key]).all(), key
    assert (oracle.gaussian_blur(gray, (5, 5)) == golden["gauss_gray_5_5_0"]).all()
    assert (oracle.gaussian_blur(bgra, (5, 5)) == golden["gauss_bgra_5_5_0"]).all()


def test_cv2_gray_resize_bit_exact(oracle, golden):
    bgr, _, _, _ = _inputs(oracle, golden)
    assert (oracle.bgr_to_gray(bgr) == golden["bgr2gray"]).all()
    for key in golden.files:
        if key.startswith("resize_bgr_"):
            dr, dc = [int(x) for x in key[len("resize_bgr_"):].split("_")]
            assert (oracle.resize_bilinear(bgr, dr, dc) == golden[key]).all(), key
    big = oracle.fill_u8(11, 64 * 96 * 3).reshape(64, 96, 3)
    assert (oracle.resize_bilinear(big, 16, 24) == golden["resize4x_bgr"]).all()


def test_cv2_f32_within_tolerance(oracle, golden):
    _, _, f, _ = _inputs(oracle, golden)
    s = oracle.sobel3(f, ("gx", "gy", "mag"))
    assert np.abs(s["gx"] - golden["sobel_gx"]).max() < 2e-6
    assert np.abs(s["gy"] - golden["sobel_gy"]).max() < 2e-6
    assert np.abs(s["mag"] - golden["sobel_mag"]).max() < 4e-6
    H, W = f.shape
    M = oracle.rotation_matrix((W - 1) / 2, (H - 1) / 2, 15.0)
    assert np.abs(M.reshape(2, 3) - golden["rotM"]).max() < 1e-12
    w = oracle.warp_affine(f, M)
    # OpenCV quantises source coordinates to 1/32 px: loose sanity check only
    assert np.abs(w - golden["warp_f32"]).max() < 0.05 and np.abs(w - golden["warp_f32"]).mean() < 0.01
    assert np.abs(oracle.resize_bilinear(f, 30, 41) - golden["resize_f32_30_41"]).max() < 1e-5
    assert np.abs(oracle.gaussian_blur(f, (5, 5), 1.1, 1.1) - golden["gauss_f32_5_5_1.1"]).max() < 1e-6
    lap = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    assert np.abs(oracle.filter2d(f, lap) - golden["filter2d_f32_lap"]).max() < 1e-5


def test_cv2_filter2d_u8(oracle, golden):
    bgr, _, _, _ = _inputs(oracle, golden)
    lap = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    assert (oracle.filter2d(bgr, lap / 3 + 0.2) == golden["filter2d_bgr"]).all()


# ---- (d) strides and degenerate shapes ------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (2, 2), (3, 5), (5, 3)])
def test_tiny_images_reflect101(oracle, shape):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    got = oracle.gaussian_blur(img, (5, 5))
    # brute force with numpy pad(mode="reflect") semantics generalised to len < radius
    def refl(p, n):
        if n == 1:
            return 0
        while p < 0 or p >= n:
            p = -p if p < 0 else 2 * (n - 1) - p
        return p
    k = np.array([1, 4, 6, 4, 1])
    H, W = shape
    want = np.zeros_like(img)
    for y in range(H):
        for x in range(W):
            acc = np.zeros(3, np.int64)
            for i in range(5):
                for j in range(5):
                    acc += k[i] * k[j] * img[refl(y + i - 2, H), refl(x + j - 2, W)].astype(np.int64)
            want[y, x] = (acc + 128) >> 8
    assert (got == want).all()


def test_strided_inputs_equal_packed(oracle):
    img = oracle.fill_u8(13, 33 * 47 * 3).reshape(33, 47, 3)
    p = oracle.padded(img, 256)
    assert (oracle.gaussian_blur(p, (5, 5)) == oracle.gaussian_blur(img, (5, 5))).all()
    assert (oracle.bgr_to_gray(p) == oracle.bgr_to_gray(img)).all()
    assert (oracle.resize_bilinear(p, 20, 31) == oracle.resize_bilinear(img, 20, 31)).all()


def test_convert_to_model(oracle):
    rng = np.random.default_rng(3)
    u = rng.integers(0, 256, size=(17, 29, 3), dtype=np.uint8)
    f = oracle.convert_to(u, np.float32, 1.0 / 255.0, 0.0)
    assert f.dtype == np.float32 and (f == (u.astype(np.float32) * np.float32(1.0 / 255.0))).all()
    back = oracle.convert_to(f, np.uint8, 255.0, 0.0)
    assert (back == u).all()
    # half-to-even rounding and saturation
    v = np.array([[0.5, 1.5, 2.5, -3.0, 254.5, 255.5, 300.0, np.float32(1e9)]], np.float32)
    assert oracle.convert_to(v, np.uint8).ravel().tolist() == [0, 2, 2, 0, 254, 255, 255, 255]
    cv2 = pytest.importorskip("cv2")
    assert np.abs(cv2.convertScaleAbs(v) .astype(int) - oracle.convert_to(np.abs(v), np.uint8).astype(int)).max() <= 1
    assert (oracle.convert_to(u, np.float32) == u.astype(np.float32)).all()


@pytest.mark.parametrize("cn", [1, 3, 4])
def test_resize_closed_forms_at_exact_scales(oracle, cn):
    """The closed forms the fast-path kernels implement (k_resize2x_u8, k_resize_up2x_u8, k_resize4x_u8c3) ARE
    the oracle's fixed-point model at scales 2, 1/2 and 4 -- and OpenCV's, where cv2 is installed."""
    rng = np.random.default_rng(40 + cn)
    h, w = 24, 36
    a = rng.integers(0, 256, size=(h, w, cn), dtype=np.uint8)
    a3 = a.astype(np.int64)
    img = a.reshape(h, w) if cn == 1 else a
    # 2x down: rounded 2x2 box mean
    box2 = ((a3.reshape(h // 2, 2, w // 2, 2, cn).sum(axis=(1, 3)) + 2) >> 2).astype(np.uint8)
    down2 = oracle.resize_bilinear(img, h // 2, w // 2).reshape(h // 2, w // 2, cn)
    assert (down2 == box2).all()
    # 4x down: the middle 2x2 of every 4x4
    mid = a3.reshape(h // 4, 4, w // 4, 4, cn)[:, 1:3, :, 1:3, :]
    box4 = ((mid.sum(axis=(1, 3)) + 2) >> 2).astype(np.uint8)
    assert (oracle.resize_bilinear(img, h // 4, w // 4).reshape(h // 4, w // 4, cn) == box4).all()
    # 2x up: H = c[far] + 3 c[near]; out = ((H_far >> 2) + ((3 H_near) >> 2) + 2) >> 2, far taps clamped
    dx, dy = np.arange(2 * w), np.arange(2 * h)
    far = np.where(dx % 2 == 0, np.maximum(dx // 2 - 1, 0), np.minimum(dx // 2 + 1, w - 1))
    farr = np.where(dy % 2 == 0, np.maximum(dy // 2 - 1, 0), np.minimum(dy // 2 + 1, h - 1))
    H = a3[:, far, :] + 3 * a3[:, dx // 2, :]
    up2 = (((H[farr] >> 2) + ((3 * H[dy // 2]) >> 2) + 2) >> 2).astype(np.uint8)
    assert (oracle.resize_bilinear(img, 2 * h, 2 * w).reshape(2 * h, 2 * w, cn) == up2).all()
    cv2 = pytest.importorskip("cv2")
    assert (cv2.resize(img, (w // 2, h // 2), interpolation=cv2.INTER_LINEAR).reshape(box2.shape) == box2).all()
    assert (cv2.resize(img, (2 * w, 2 * h), interpolation=cv2.INTER_LINEAR).reshape(up2.shape) == up2).all()


def test_gaussian3_binomial_closed_form(oracle):
    """Gauss3Op's arithmetic: with Q8 taps {64,128,64} the oracle's (sum + 2^15) >> 16 is (S + 8) >> 4 of the
    {1,2,1} x {1,2,1} sum S, which the kernel forms as the high byte of 16 S + 128."""
    assert oracle.gaussian_kernel_q8(3, 0.0).tolist() == [64, 128, 64]
    rng = np.random.default_rng(44)
    img = rng.integers(0, 256, size=(19, 23), dtype=np.uint8)
    p = np.pad(img.astype(np.int64), 1, mode="reflect")  # numpy "reflect" == BORDER_REFLECT_101
    b = np.array([1, 2, 1])
    S = sum(b[i] * b[j] * p[i:i + 19, j:j + 23] for i in range(3) for j in range(3))
    want = oracle.gaussian_blur(img, (3, 3))
    assert (((S + 8) >> 4) == want).all()
    assert ((((16 * S + 128) >> 8) & 0xFF) == want).all() and (16 * S + 128).max() <= 65408


def test_fused_sobel_chain_is_integer_exact(oracle):
    """YuyvSobelOp keeps the Sobel sums of the gray image in integer registers: on u8-valued f32 input every
    intermediate of orc_sobel3_f32 is an integer below 2^24, so the f32 chain and the integer chain agree and only
    the final sqrtf rounds."""
    rng = np.random.default_rng(45)
    g = rng.integers(0, 256, size=(31, 37)).astype(np.int64)
    g[3:9, 4:20] = 255
    g[10:14, 0:9] = 0
    out = oracle.sobel3(g.astype(np.float32), want=("gx", "gy", "mag"))
    p = np.pad(g, 1, mode="reflect")
    gx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    gy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    assert (out["gx"] == gx).all() and (out["gy"] == gy).all()
    ss = gx * gx + gy * gy
    assert ss.max() < 2 ** 24
    assert (out["mag"] == np.sqrt(ss.astype(np.float32))).all()  # numpy's f32 sqrt is correctly rounded too


def test_mjpeg_fixture_is_what_libjpeg_turbo_decodes():
    """tests/golden/mjpeg_golden.npz: the stored BGR frames are cv2.imdecode (libjpeg-turbo) of the stored JPEG
    bytes -- the pin of the MJPEG branch (skipped where cv2 is absent, e.g. on the GPU box)."""
    import os

    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mjpeg_golden.npz"))
    names = [k[5:] for k in g.files if k.startswith("jpeg_")]
    assert len(names) >= 6
    for n in names:
        assert (cv2.imdecode(g[f"jpeg_{n}"], cv2.IMREAD_COLOR) == g[f"bgr_{n}"]).all(), n


def test_mjpeg_colour_conversion_formula_is_libjpeg_turbos():
    """k_ycc_to_bgr's fixed-point YCbCr -> BGR (csrc/mjpeg.cu: 91881 / 116130 / 22554 / 46802, ONE_HALF, arithmetic
    shifts, clamp) applied to libjpeg's own full-resolution YCbCr output (Pillow, JCS_YCbCr) reproduces
    libjpeg-turbo's BGR (cv2.imdecode) bit for bit on every fixture -- the CPU-side pin of that kernel's arithmetic;
    its upsampling half is pinned on the GPU against the same BGR frames."""
    import io
    import os

    cv2 = pytest.importorskip("cv2")
    Image = pytest.importorskip("PIL.Image")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mjpeg_golden.npz"))
    for n in ("444_q95", "420_q90", "422_q85", "420_odd_q92", "422_odd_q92"):
        im = Image.open(io.BytesIO(g[f"jpeg_{n}"].tobytes()))
        im.draft("YCbCr", im.size)
        assert im.mode == "YCbCr"
        ycc = np.asarray(im).astype(np.int64)
        y, cb, cr = ycc[..., 0], ycc[..., 1] - 128, ycc[..., 2] - 128
        r = y + ((91881 * cr + 32768) >> 16)
        b = y + ((116130 * cb + 32768) >> 16)
        gg = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
        bgr = np.clip(np.stack([b, gg, r], -1), 0, 255).astype(np.uint8)
        assert (bgr == g[f"bgr_{n}"]).all(), n
        assert (bgr == cv2.imdecode(g[f"jpeg_{n}"], cv2.IMREAD_COLOR)).all(), n


def test_tuned_cpu_gaussian_is_bit_identical_to_the_definition(oracle):
    """oracle.gaussian5_fast (the auto-vectorised restatement bench.py times as its second CPU figure) equals the
    definition on every geometry class, and reproduces the 4K golden CRC of SURVEY.md section 8c."""
    for shape in ((5, 5, 3), (7, 9, 1), (64, 83, 3), (33, 40, 4), (1, 1, 3), (2, 3, 3), (3, 1, 2), (120, 301, 3)):
        a = oracle.fill_u8(1 + shape[0], int(np.prod(shape))).reshape(shape)
        if shape[2] == 1:
            a = a.reshape(shape[:2])
        assert (oracle.gaussian5_fast(a) == oracle.gaussian_blur(a, (5, 5))).all(), shape
    p = oracle.padded(oracle.fill_u8(13, 33 * 47 * 3).reshape(33, 47, 3), 256)
    assert (oracle.gaussian5_fast(p) == oracle.gaussian_blur(p, (5, 5))).all()
    img = oracle.fill_u8(2, 2160 * 3840 * 3).reshape(2160, 3840, 3)
    assert oracle.crc32(oracle.gaussian5_fast(img)) == 0x827081C8


def test_gray_pair_byte_split_identity_exhaustive():
    """cvt_math.cuh gray_pair: with 3735 = 14*256+151, 19235 = 75*256+35, 9798 = 38*256+70,
    (3735b + 19235g + 9798r + 16384) >> 15 == (A + (B >> 8) + 64) >> 7 for every (b, g, r), and A, B fit 16 bits."""
    b, g, r = np.meshgrid(np.arange(256, dtype=np.int64), np.arange(256, dtype=np.int64), np.arange(256, dtype=np.int64), indexing="ij")
    A = 14 * b + 75 * g + 38 * r
    B = 151 * b + 35 * g + 70 * r
    assert A.max() <= 32385 and B.max() <= 65280 and (A + (B >> 8) + 64).max() < 65536
    assert (((3735 * b + 19235 * g + 9798 * r + 16384) >> 15) == ((A + (B >> 8) + 64) >> 7)).all()


def test_transposed_form_recurrences_equal_the_windowed_sums():
    """The strip ops keep partial sums instead of a window of rows (Gauss5Op, GaussQ8Op, the dense filter ops):
    s4 = x_n, s3 = x_{n-1} + 4 x_n, s2 = x_{n-2} + 4 x_{n-1} + 6 x_n, s1 = x_{n-3} + 4 x_{n-2} + 6 x_{n-1} + 4 x_n and
    V = s1 + x_{n+1}; in general out = s_0 + K[KS-1] x, s_i = s_{i+1} + K[KS-2-i] x, s_{KS-2} = K[0] x.  Checked in
    integers against the direct vertical sums for the binomial 5 taps and random symmetric Q8 taps of 3..15, whatever
    the state held before the first KS-1 rows (the ops never reset it)."""
    rng = np.random.default_rng(11)
    for ks in (3, 5, 7, 9, 15):
        taps = [np.array([1, 4, 6, 4, 1])] if ks == 5 else []
        for _ in range(3):
            h = rng.integers(0, 60, size=ks // 2 + 1)
            k = np.concatenate([h, h[-2::-1]])
            taps.append(k)
        for K in taps:
            rows = rng.integers(0, 256, size=(40, 16)).astype(np.int64)
            state = [rng.integers(0, 1000, size=16).astype(np.int64) for _ in range(ks - 1)]  # garbage on purpose
            outs = []
            for x in rows:
                outs.append(state[0] + K[ks - 1] * x)
                for i in range(ks - 2):
                    state[i] = state[i + 1] + K[ks - 2 - i] * x
                state[ks - 2] = K[0] * x
            for n in range(ks - 1, len(rows)):
                want = sum(K[i] * rows[n - (ks - 1) + i] for i in range(ks))
                assert (outs[n] == want).all(), (ks, n)


def test_dense_transposed_form_keeps_the_row_major_f32_order():
    """Filter2dF32CnOp / Filter2dU8Op: every pending output row is a running fmaf chain that meets its kernel rows in
    ky order and, inside a row, its taps in kx order -- the oracle's row-major order, so the f32 result is bit-identical.
    The two sides use the same fma emulation: what is compared is the ORDER of the chain, not the rounding model."""
    rng = np.random.default_rng(12)

    def fmaf(a, b, c):
        return np.float32(np.float64(a) * np.float64(b) + np.float64(c))

    ks = 5
    k = rng.normal(size=(ks, ks)).astype(np.float32)
    img = rng.random(size=(12, 12)).astype(np.float32)
    delta = np.float32(0.25)
    # oracle order for the one output whose window is img[0:5, 3:8]
    acc = delta
    for i in range(ks):
        for j in range(ks):
            acc = fmaf(k[i, j], img[i, 3 + j], acc)
    # transposed: the pending sum of that output row as the rows 0..4 arrive
    pend = None
    for r in range(ks):
        s = delta if r == 0 else pend
        for j in range(ks):
            s = fmaf(k[r, j], img[r, 3 + j], s)
        pend = s
    assert pend.view(np.int32) == acc.view(np.int32)
