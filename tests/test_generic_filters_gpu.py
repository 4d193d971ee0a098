"""GPU suite, part 3: the general-case (register-tiled) separable / dense filter kernels of filter.cu -- every tap
count (compile-time 3 / 5 / 7 and run-time chains incl. 1, even sizes, 9..31, non-square), every channel count,
u8 and f32, images that span several CTA tiles with ragged edges, batches.  Bit-exact against the oracle (f32:
the oracle's fmaf order, 0 ULP)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORCE = ("gauss.force_generic", "sepf32.force_generic", "f2d.force_generic")


@pytest.fixture()
def generic(rcv):
    for o in FORCE:
        rcv.imgproc.set_option(o, 1)
    yield rcv
    for o in FORCE:
        rcv.imgproc.set_option(o, 0)


def _img(oracle, seed, h, w, cn, f32):
    shape = (h, w) if cn == 1 else (h, w, cn)
    if f32:
        return oracle.fill_f32(seed, h * w * cn).reshape(shape)
    return oracle.fill_u8(seed, h * w * cn).reshape(shape)


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("kw,kh", [(1, 1), (2, 2), (3, 3), (5, 5), (7, 7), (9, 9), (11, 11), (15, 15), (31, 31), (9, 3), (3, 13),
                                   (8, 5), (16, 17)])
def test_separable_u8_q8_general_kernel(generic, oracle, cn, kw, kh):
    R = generic
    rng = np.random.default_rng(kw * 100 + kh)
    h, w = 75, 301  # 3 tile rows of 32; 301*cn element columns: 2..5 tiles of 240/256 with a ragged last one
    a = _img(oracle, 2000 + cn, h, w, cn, False)
    kx = rng.integers(-40, 90, size=kw).astype(np.int32)
    ky = rng.integers(-40, 90, size=kh).astype(np.int32)
    s = R.Mat.from_numpy(a).upload()
    d = s.like()
    R.imgproc.sep_filter2d(s, d, kx, ky)
    assert (d.to_numpy() == oracle.sepfilter_u8_q8(a, kx, ky)).all(), f"cn{cn} {kw}x{kh}"


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("kw,kh", [(1, 1), (3, 3), (5, 5), (7, 7), (9, 9), (13, 13), (31, 31), (5, 11), (12, 4)])
def test_separable_f32_general_kernel(generic, oracle, cn, kw, kh):
    R = generic
    rng = np.random.default_rng(kw * 100 + kh + 7)
    h, w = 70, 203
    a = _img(oracle, 2100 + cn, h, w, cn, True)
    kx = rng.normal(size=kw).astype(np.float32)
    ky = rng.normal(size=kh).astype(np.float32)
    s = R.Mat.from_numpy(a).upload()
    d = s.like()
    R.imgproc.sep_filter2d(s, d, kx, ky)
    want = oracle.sepfilter_f32(a, kx, ky)
    assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"cn{cn} {kw}x{kh}"


@pytest.mark.parametrize("cn", [1, 3, 4])
@pytest.mark.parametrize("kw,kh", [(1, 1), (3, 3), (5, 5), (7, 7), (9, 9), (4, 4), (7, 3), (3, 11), (15, 15)])
@pytest.mark.parametrize("f32", [False, True])
def test_dense_filter2d_general_kernel(generic, oracle, cn, kw, kh, f32):
    R = generic
    rng = np.random.default_rng(kw * 31 + kh)
    h, w = 53, 270  # 4 tile rows of 16; ragged tiles
    a = _img(oracle, 2200 + cn, h, w, cn, f32)
    k = rng.normal(size=(kh, kw)).astype(np.float32)
    if not f32:
        k = (k / max(1e-3, np.abs(k).sum()) * 1.7).astype(np.float32)  # some saturation, both ways
    s = R.Mat.from_numpy(a).upload()
    d = s.like()
    R.imgproc.filter2d(s, d, k, delta=0.375)
    want = oracle.filter2d(a, k, 0.375)
    if f32:
        assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"cn{cn} {kw}x{kh}"
    else:
        assert (d.to_numpy() == want).all(), f"cn{cn} {kw}x{kh}"


def test_general_kernels_tiny_images_and_batches(generic, oracle):
    R = generic
    for shape in ((1, 1), (1, 9), (9, 1), (2, 3), (5, 4), (33, 2)):
        a = _img(oracle, 2300, shape[0], shape[1], 3, False)
        d = R.Mat.empty()
        R.imgproc.gaussian_blur(R.Mat.from_numpy(a), d, (9, 9), 2.0)
        assert (d.to_numpy() == oracle.gaussian_blur(a, (9, 9), 2.0, 2.0)).all(), shape
        f = _img(oracle, 2301, shape[0], shape[1], 1, True)
        k = np.arange(25, dtype=np.float32).reshape(5, 5) / 25
        R.imgproc.filter2d(R.Mat.from_numpy(f), d, k)
        assert (d.to_numpy().view(np.int32) == oracle.filter2d(f, k, 0.0).view(np.int32)).all(), shape
    # a batch of device frames: one launch, grid.z = frames
    n, h, w = 3, 40, 500
    frames = [_img(oracle, 2400 + j, h, w, 3, False) for j in range(n)]
    src, dst = R.Mat.device_batch(n, h, w, 3), R.Mat.device_batch(n, h, w, 3)
    import ctypes as C
    from rustcv_b200 import _ffi as F
    for j in range(n):
        hm = R.Mat.from_numpy(frames[j])
        F.check(F.lib.rcv_mat_upload(C.byref(hm.c()), C.byref(src[j].c())))
    n0 = R.imgproc.launch_count()
    R.imgproc.gaussian_blur_batch(src, dst, (13, 13), 2.0, 2.0)
    assert R.imgproc.launch_count() - n0 == 1
    for j in range(n):
        assert (dst[j].to_numpy() == oracle.gaussian_blur(frames[j], (13, 13), 2.0, 2.0)).all(), f"frame {j}"


# ---- multi-channel f32 strip ops (strip_f32cn.cu): several halo lanes per side ----------------------------------------
@pytest.mark.parametrize("cn", [2, 3, 4])
@pytest.mark.parametrize("ks", [3, 5, 7])
def test_f32_multichannel_strip_kernels(rcv, oracle, cn, ks):
    """k_strip<SepF32CnOp<KS,CN>> / k_strip<Filter2dF32CnOp<3,CN>>: f32 BGR / BGRA / 2-channel, several strips with a
    ragged last one, band seams, borders; 0 ULP vs the oracle; the general kernel agrees."""
    R = rcv
    rng = np.random.default_rng(ks * 10 + cn)
    for (h, w) in ((203, 517), (64, 16), (9, 40)):
        a = oracle.fill_f32(3000 + ks + cn, h * w * cn).reshape(h, w, cn)
        s = R.Mat.from_numpy(a).upload()
        d = s.like()
        kx = rng.normal(size=ks).astype(np.float32)
        ky = rng.normal(size=ks).astype(np.float32)
        n0 = R.imgproc.launch_count()
        R.imgproc.sep_filter2d(s, d, kx, ky)
        assert R.imgproc.launch_count() - n0 == 1
        want = oracle.sepfilter_f32(a, kx, ky)
        assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"sep cn{cn} ks{ks} {h}x{w}"
        R.imgproc.gaussian_blur(s, d, (ks, ks), 1.3)
        wg = oracle.gaussian_blur(a, (ks, ks), 1.3, 1.3)
        assert (d.to_numpy().view(np.int32) == wg.view(np.int32)).all(), f"gauss cn{cn} ks{ks} {h}x{w}"
        if True:  # dense filter2D: the transposed-form strip op covers every (ks, cn) here
            k = rng.normal(size=(ks, ks)).astype(np.float32)
            R.imgproc.filter2d(s, d, k, delta=-0.25)
            wf = oracle.filter2d(a, k, -0.25)
            assert (d.to_numpy().view(np.int32) == wf.view(np.int32)).all(), f"filter2d cn{cn} {h}x{w}"
    # band seams
    a = oracle.fill_f32(3100 + ks + cn, 150 * 300 * cn).reshape(150, 300, cn)
    s = R.Mat.from_numpy(a).upload()
    kx = rng.normal(size=ks).astype(np.float32)
    want = oracle.sepfilter_f32(a, kx, kx)
    for br in (8, 12, 36):
        R.imgproc.set_option("sepf32.band_rows", br)
        d = s.like()
        R.imgproc.sep_filter2d(s, d, kx, kx)
        assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"band_rows {br}"
    R.imgproc.set_option("sepf32.band_rows", 0)
    # host Mats (banded pinned pipeline uses the row-window form of the op)
    big = oracle.fill_f32(3200 + cn, 700 * 900 * cn).reshape(700, 900, cn)
    hp = R.Mat.pinned(700, 900, cn, R.F32)
    hp.data[:] = big.view(np.uint8).ravel()
    hd = R.Mat.pinned(700, 900, cn, R.F32)
    R.imgproc.sep_filter2d(hp, hd, kx, kx)
    assert (hd.to_numpy().view(np.int32) == oracle.sepfilter_f32(big, kx, kx).view(np.int32)).all(), "pinned host, banded"


# ---- wide u8 Gaussians in the strip pipeline (strip_gaussq8_wide.cuh): 9..15 taps, 16-row chunks, 8 warps per CTA -------
@pytest.mark.parametrize("cn", [1, 3, 4])
@pytest.mark.parametrize("ks", [9, 11, 13, 15])
def test_gaussian_wide_strip_kernel(rcv, oracle, cn, ks):
    R = rcv
    for (h, w) in ((203, 517), (16, 16), (40, 700), (130, 161)):
        a = oracle.fill_u8(4000 + ks + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
        s = R.Mat.from_numpy(a).upload()
        d = s.like()
        for sx, sy in ((0.0, 0.0), (2.0, 2.0), (1.1, 3.7)):
            n0 = R.imgproc.launch_count()
            R.imgproc.gaussian_blur(s, d, (ks, ks), sx, sy)
            assert R.imgproc.launch_count() - n0 == 1
            want = oracle.gaussian_blur(a, (ks, ks), sx, sy)
            got = d.to_numpy()
            assert (got == want).all(), (f"cn{cn} ks{ks} {h}x{w} sigma {sx},{sy}: {(got != want).sum()} differ, first at "
                                         f"{np.argwhere(got != want)[:3].tolist()}")
    # band seams, saturated regions, a batch in one launch, host Mats (banded pinned pipeline: row windows)
    h, w = 300, 420
    a = oracle.fill_u8(4100 + ks + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
    a[40:90, 100:300] = 255
    a[200:240, :64] = 0
    want = oracle.gaussian_blur(a, (ks, ks), 2.5, 2.5)
    s = R.Mat.from_numpy(a).upload()
    for br in (8, 24, 50):
        R.imgproc.set_option("gauss.band_rows", br)
        d = s.like()
        R.imgproc.gaussian_blur(s, d, (ks, ks), 2.5)
        assert (d.to_numpy() == want).all(), f"band_rows {br}"
    R.imgproc.set_option("gauss.band_rows", 0)
    big = oracle.fill_u8(4200 + cn, 1200 * 1000 * cn).reshape((1200, 1000) if cn == 1 else (1200, 1000, cn))
    hp = R.Mat.pinned(1200, 1000, cn)
    hp.data[:] = big.ravel()
    hd = R.Mat.pinned(1200, 1000, cn)
    R.imgproc.gaussian_blur(hp, hd, (ks, ks), 2.0)
    assert (hd.to_numpy() == oracle.gaussian_blur(big, (ks, ks), 2.0, 2.0)).all(), "pinned host Mat (banded)"


def test_gaussian_wide_falls_back_for_tiny_or_two_channel_images(rcv, oracle):
    R = rcv
    for shape in ((9, 40, 3), (40, 9, 3), (64, 64, 2)):
        a = oracle.fill_u8(4300, int(np.prod(shape))).reshape(shape)
        d = R.Mat.empty()
        R.imgproc.gaussian_blur(R.Mat.from_numpy(a), d, (11, 11), 2.0)
        assert (d.to_numpy() == oracle.gaussian_blur(a, (11, 11), 2.0, 2.0)).all(), shape


@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("ks", [9, 11])
def test_gaussian_9_and_11_taps_both_strip_ops_agree(rcv, oracle, cn, ks):
    """9 / 11 taps run the horizontal-first op (GaussQ8Op, 8 warps, 16-row chunks) where the taps stay within the
    adjacent lanes, the windowed wide op otherwise or on request: both bit-exact, several band heights, ragged strips."""
    R = rcv
    for (h, w) in ((203, 517), (16, 16), (130, 1000)):
        a = oracle.fill_u8(4400 + ks + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
        want = oracle.gaussian_blur(a, (ks, ks), 1.9, 2.3)
        s = R.Mat.from_numpy(a).upload()
        for windowed in (0, 1):
            for br in (0, 16, 40):
                R.imgproc.set_option("gauss.wide_windowed", windowed)
                R.imgproc.set_option("gauss.band_rows", br)
                try:
                    d = s.like()
                    R.imgproc.gaussian_blur(s, d, (ks, ks), 1.9, 2.3)
                finally:
                    R.imgproc.set_option("gauss.wide_windowed", 0)
                    R.imgproc.set_option("gauss.band_rows", 0)
                assert (d.to_numpy() == want).all(), f"cn{cn} ks{ks} {h}x{w} windowed{windowed} band{br}"


@pytest.mark.parametrize("cn", [1, 3])
@pytest.mark.parametrize("ks", [9, 11, 13, 15])
def test_f32_wide_strip_kernel(rcv, oracle, cn, ks):
    """k_strip<SepF32WideOp<KS,CN>>: 9..15 taps on gray / BGR f32 (16-row chunks, up to 6 halo lanes per side):
    0 ULP vs the oracle on several strips, ragged edges, borders, band seams, pinned host Mats."""
    R = rcv
    rng = np.random.default_rng(ks * 7 + cn)
    for (h, w) in ((203, 517), (16, 32), (40, 300)):
        a = oracle.fill_f32(5000 + ks + cn, h * w * cn).reshape((h, w) if cn == 1 else (h, w, cn))
        s = R.Mat.from_numpy(a).upload()
        d = s.like()
        kx = rng.normal(size=ks).astype(np.float32)
        ky = rng.normal(size=ks).astype(np.float32)
        n0 = R.imgproc.launch_count()
        R.imgproc.sep_filter2d(s, d, kx, ky)
        assert R.imgproc.launch_count() - n0 == 1
        want = oracle.sepfilter_f32(a, kx, ky)
        assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"sep cn{cn} ks{ks} {h}x{w}"
        R.imgproc.gaussian_blur(s, d, (ks, ks), 2.1)
        wg = oracle.gaussian_blur(a, (ks, ks), 2.1, 2.1)
        assert (d.to_numpy().view(np.int32) == wg.view(np.int32)).all(), f"gauss cn{cn} ks{ks} {h}x{w}"
    a = oracle.fill_f32(5100 + ks + cn, 150 * 300 * cn).reshape((150, 300) if cn == 1 else (150, 300, cn))
    s = R.Mat.from_numpy(a).upload()
    kx = rng.normal(size=ks).astype(np.float32)
    want = oracle.sepfilter_f32(a, kx, kx)
    for br in (8, 30, 50):
        R.imgproc.set_option("sepf32.band_rows", br)
        d = s.like()
        R.imgproc.sep_filter2d(s, d, kx, kx)
        assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"band_rows {br}"
    R.imgproc.set_option("sepf32.band_rows", 0)
    big = oracle.fill_f32(5200 + cn, 700 * 900 * cn).reshape((700, 900) if cn == 1 else (700, 900, cn))
    hp = R.Mat.pinned(700, 900, cn, R.F32)
    hp.data[:] = big.view(np.uint8).ravel()
    hd = R.Mat.pinned(700, 900, cn, R.F32)
    R.imgproc.sep_filter2d(hp, hd, kx, kx)
    assert (hd.to_numpy().view(np.int32) == oracle.sepfilter_f32(big, kx, kx).view(np.int32)).all(), "pinned host, banded"


def test_round2_strip_ops_random_geometries(rcv, oracle):
    """40 random (rows, cols, channels, taps, band height, location) draws over the strip ops added in round 2 --
    multi-channel f32 (3 / 5 / 7 taps), wide u8 and wide f32 (9..15 taps): every ragged-edge / band-seam / partial-chunk
    / halo-lane combination they can meet, against the oracle (bit-exact / 0 ULP)."""
    R = rcv
    rng = np.random.default_rng(20261019)

    def place(a, where):
        if where == "host":
            return R.Mat.from_numpy(a), R.Mat.empty()
        s = R.Mat.from_numpy(a).upload()
        return s, s.like()

    for it in range(40):
        h = int(rng.integers(16, 300))
        w = int(rng.integers(32, 700))
        band = int(rng.choice([0, 8, 12, 24, 40, 98]))
        where = str(rng.choice(["device", "host"]))
        R.imgproc.set_option("gauss.band_rows", band)
        R.imgproc.set_option("sepf32.band_rows", band)
        try:
            # wide u8 Gaussian
            cn = int(rng.choice([1, 3, 4]))
            ks = int(rng.choice([9, 11, 13, 15]))
            sx, sy = float(rng.uniform(0.8, 4.0)), float(rng.uniform(0.8, 4.0))
            a = rng.integers(0, 256, size=(h, w, cn), dtype=np.uint8)
            if cn == 1:
                a = a.reshape(h, w)
            s, d = place(a, where)
            R.imgproc.gaussian_blur(s, d, (ks, ks), sx, sy)
            assert (d.to_numpy() == oracle.gaussian_blur(a, (ks, ks), sx, sy)).all(), f"wide u8 it{it} {h}x{w}x{cn} ks{ks} band{band} {where}"
            # multi-channel f32, 3 / 5 / 7 taps
            cn = int(rng.integers(2, 5))
            ks = int(rng.choice([3, 5, 7]))
            f = rng.random(size=(h, w, cn), dtype=np.float32)
            kx = rng.normal(size=ks).astype(np.float32)
            ky = rng.normal(size=ks).astype(np.float32)
            s, d = place(f, where)
            R.imgproc.sep_filter2d(s, d, kx, ky)
            assert (d.to_numpy().view(np.int32) == oracle.sepfilter_f32(f, kx, ky).view(np.int32)).all(), f"f32 cn it{it} {h}x{w}x{cn} ks{ks} band{band} {where}"
            # wide f32, gray / BGR
            cn = int(rng.choice([1, 3]))
            ks = int(rng.choice([9, 11, 13, 15]))
            f = rng.random(size=(h, w, cn), dtype=np.float32)
            if cn == 1:
                f = f.reshape(h, w)
            kx = rng.normal(size=ks).astype(np.float32)
            s, d = place(f, where)
            R.imgproc.sep_filter2d(s, d, kx, kx)
            assert (d.to_numpy().view(np.int32) == oracle.sepfilter_f32(f, kx, kx).view(np.int32)).all(), f"wide f32 it{it} {h}x{w}x{cn} ks{ks} band{band} {where}"
        finally:
            R.imgproc.set_option("gauss.band_rows", 0)
            R.imgproc.set_option("sepf32.band_rows", 0)


# ---- dense filter2D strip ops in transposed form (strip_f32cn.cu Filter2dF32CnOp, strip_f2d_u8.cu Filter2dU8Op) -------
@pytest.mark.parametrize("cn", [1, 2, 3, 4])
@pytest.mark.parametrize("ks", [3, 5, 7])
def test_dense_filter2d_strip_kernels_transposed_form(rcv, oracle, cn, ks):
    """Pending-row-sum form of the dense filters: f32 3x3 / 5x5 / 7x7 on 1..4 channels (0 ULP vs the oracle -- the
    accumulation order must be the oracle's row-major one) and u8 3x3 / 5x5 / 7x7 (bit-exact); an asymmetric ramp kernel
    catches any swap of rows, columns or direction; several strips with a ragged last one, forced band heights down to
    one chunk (warm-up rows hoisted out of the steady loop), one launch per call, host Mats through the banded
    pipeline, and agreement with the general kernel."""
    R = rcv
    rng = np.random.default_rng(ks * 100 + cn)
    ramp = (np.arange(ks * ks, dtype=np.float32).reshape(ks, ks) - 3.0) / (ks * ks * 4)
    for (h, w) in ((203, 517), (64, 16), (9, 40), (150, 300)):
        shp = (h, w) if cn == 1 else (h, w, cn)
        a = oracle.fill_f32(5000 + ks + cn, h * w * cn).reshape(shp)
        s = R.Mat.from_numpy(a).upload()
        d = s.like()
        for k, delta in ((ramp, 0.0), (rng.normal(size=(ks, ks)).astype(np.float32), -0.25)):
            for br in (0, 8, 20):
                R.imgproc.set_option("f2d.band_rows", br)
                try:
                    n0 = R.imgproc.launch_count()
                    R.imgproc.filter2d(s, d, k, delta=delta)
                    assert R.imgproc.launch_count() - n0 == 1
                finally:
                    R.imgproc.set_option("f2d.band_rows", 0)
                want = oracle.filter2d(a, k, delta)
                assert (d.to_numpy().view(np.int32) == want.view(np.int32)).all(), f"f32 cn{cn} ks{ks} {h}x{w} band{br}"
        if True:  # u8: 3x3 / 5x5 / 7x7 (7x7: 12 warps per CTA, three halo words per side)
            u = oracle.fill_u8(5100 + ks + cn, h * w * cn).reshape(shp)
            su = R.Mat.from_numpy(u).upload()
            du = su.like()
            ku = (ramp * 3).astype(np.float32)
            for br in (0, 8, 20):
                R.imgproc.set_option("f2d.band_rows", br)
                try:
                    R.imgproc.filter2d(su, du, ku, delta=0.5)
                finally:
                    R.imgproc.set_option("f2d.band_rows", 0)
                assert (du.to_numpy() == oracle.filter2d(u, ku, 0.5)).all(), f"u8 cn{cn} ks{ks} {h}x{w} band{br}"
    # the general kernel agrees (same inputs, strip ops off)
    R.imgproc.set_option("f2d.force_generic", 1)
    try:
        d2 = s.like()
        R.imgproc.filter2d(s, d2, ramp, delta=0.0)
    finally:
        R.imgproc.set_option("f2d.force_generic", 0)
    R.imgproc.filter2d(s, d, ramp, delta=0.0)
    assert (d2.to_numpy().view(np.int32) == d.to_numpy().view(np.int32)).all()
    # host Mats: the banded pinned pipeline runs the op on row windows
    big = oracle.fill_f32(5200 + cn, 700 * 900 * cn).reshape((700, 900) if cn == 1 else (700, 900, cn))
    hp = R.Mat.pinned(700, 900, cn, R.F32)
    hp.data[:] = big.view(np.uint8).ravel()
    hd = R.Mat.pinned(700, 900, cn, R.F32)
    R.imgproc.filter2d(hp, hd, ramp, delta=0.125)
    assert (hd.to_numpy().view(np.int32) == oracle.filter2d(big, ramp, 0.125).view(np.int32)).all(), "pinned host, banded"


def test_filter2d_batch_entry_point(rcv, oracle):
    """rcv_filter2d_batch: a uniform device batch is ONE launch (u8 5x5 and f32 7x7 strip ops, 13x13 general kernel);
    host Mats go through the staging pipeline; results equal the per-frame call's."""
    R = rcv
    import ctypes as C
    from rustcv_b200 import _ffi as F
    rng = np.random.default_rng(77)
    n, h, w = 3, 90, 400
    for depth, cn, ks in ((R.U8, 3, 5), (R.F32, 3, 7), (R.F32, 1, 3), (R.U8, 1, 13)):
        f32 = depth == R.F32
        frames = [_img(oracle, 6000 + j, h, w, cn, f32) for j in range(n)]
        src, dst = R.Mat.device_batch(n, h, w, cn, depth), R.Mat.device_batch(n, h, w, cn, depth)
        for j in range(n):
            hm = R.Mat.from_numpy(frames[j])
            F.check(F.lib.rcv_mat_upload(C.byref(hm.c()), C.byref(src[j].c())))
        k = (rng.normal(size=(ks, ks)) / ks).astype(np.float32)
        n0 = R.imgproc.launch_count()
        R.imgproc.filter2d_batch(src, dst, k, delta=0.25)
        assert R.imgproc.launch_count() - n0 == 1, "one launch for the whole device batch"
        for j in range(n):
            want = oracle.filter2d(frames[j], k, 0.25)
            got = dst[j].to_numpy()
            assert (got.view(np.int32) == want.view(np.int32)).all() if f32 else (got == want).all(), f"{depth} cn{cn} ks{ks} frame {j}"
        hs = [R.Mat.from_numpy(fr) for fr in frames]
        hd = [R.Mat.new(h, w, cn, depth) for _ in range(n)]
        R.imgproc.filter2d_batch(hs, hd, k, delta=0.25)
        for j in range(n):
            assert (hd[j].to_numpy().reshape(-1).view(np.uint8) == dst[j].to_numpy().reshape(-1).view(np.uint8)).all(), f"host batch frame {j}"
        src.free(); dst.free()
