"""CPU suite, part 2: the C-ABI boundary without a GPU.

The library must load, export exactly the symbols include/rcv_imgproc.h declares,
validate arguments before touching CUDA, and FAIL LOUDLY (never fall back to a CPU
path) when no B200 is present.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcv_imgproc.h")


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^RCV_API\s+[\w\s\*]*?\b(rcv_\w+)\s*\(", text, flags=re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    from rustcv_b200 import build

    lib_path = build.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted(set(re.findall(r" T (rcv_\w+)", out)))
    declared = _declared()
    assert len(declared) >= 35
    assert exported == declared, (set(declared) ^ set(exported))


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 and as C++ with no other includes."""
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
        r = subprocess.run(["gcc", "-x", lang, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", HEADER],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ctypes_binding_covers_the_header():
    from rustcv_b200 import _ffi

    assert sorted(_ffi.SIGNATURES) == _declared()
    assert C.sizeof(_ffi.RcvMat) == 32  # ptr, 2 x i32, size_t, 4 x u8, i32  (x86-64)


def test_no_external_cuda_library_dependency():
    from rustcv_b200 import _ffi

    out = subprocess.check_output(["ldd", _ffi.LIB_PATH], text=True)
    assert "libtorch" not in out and "libcuda.so" not in out and "libcudart" not in out  # static cudart


def test_product_does_not_import_the_oracle():
    """Nothing under rustcv_b200/ may include, import, link or dlopen oracle/ (comments
    may cite it)."""
    pkg = os.path.join(ROOT, "rustcv_b200")
    # (dlopen itself is allowed -- multi.cu resolves NCCL at run time -- but never with anything of oracle/)
    bad = re.compile(r'#\s*include\s*[<"][^>"]*oracle|^\s*(from|import)\s+[\w\.]*oracle|librcv_oracle|dlopen\([^)]*orac', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(text), f
    out = subprocess.check_output(["nm", "-D", os.path.join(pkg, "librcv_imgproc.so")], text=True)
    assert "orc_" not in out
    strings = subprocess.check_output(["strings", "-n", "6", os.path.join(pkg, "librcv_imgproc.so")], text=True)
    assert "oracle" not in strings  # no path of the checker baked into the product


def test_missing_library_fails_loudly_on_first_use():
    """The package imports (so `python -m rustcv_b200.build` works on a clean tree) but the first
    use raises ImportError -- there is no CPU fallback to fall into."""
    code = ("import rustcv_b200\n"
            "try:\n"
            "    rustcv_b200.Mat\n"
            "except ImportError as e:\n"
            "    assert 'no CPU fallback' in str(e); print('loud')\n")
    env = dict(os.environ, RCV_IMGPROC_LIB="/nonexistent/librcv_imgproc.so", PYTHONPATH=ROOT)
    out = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert out.returncode == 0 and "loud" in out.stdout, out.stderr


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu():
    import rustcv_b200 as R
    from rustcv_b200 import _ffi as F

    with pytest.raises(F.RcvError) as e:
        R.imgproc.init(0)
    assert e.value.code == F.RCV_ERR_CUDA
    src = R.Mat.from_numpy(np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(F.RcvError) as e:
        R.imgproc.gaussian_blur(src, R.Mat.empty())
    assert e.value.code == F.RCV_ERR_NOT_INIT and "no CPU fallback" in str(e.value)


def test_argument_validation_happens_before_cuda():
    import rustcv_b200 as R
    from rustcv_b200 import _ffi as F

    a = R.Mat.from_numpy(np.zeros((8, 8, 3), np.uint8))
    b = R.Mat.new(7, 8, 3)  # wrong rows; a pinned/device dst is never resized by the wrapper
    assert F.lib.rcv_gaussian_blur(C.byref(a.c()), C.byref(b.c()), 5, 5, 0.0, 0.0) == F.RCV_ERR_SIZE
    assert b"caller sizes dst" in F.lib.rcv_last_error()
    assert F.lib.rcv_gaussian_blur(None, C.byref(b.c()), 5, 5, 0.0, 0.0) == F.RCV_ERR_ARG
    c = R.Mat.new(8, 8, 1)
    assert F.lib.rcv_gaussian_blur(C.byref(a.c()), C.byref(c.c()), 5, 5, 0.0, 0.0) == F.RCV_ERR_DEPTH
    # cvtColor channel contract
    y = R.Mat.new(8, 8, 2)
    assert F.lib.rcv_cvt_color(C.byref(y.c()), C.byref(c.c()), F.COLOR_YUYV2BGR) == F.RCV_ERR_DEPTH
    assert F.lib.rcv_cvt_color(C.byref(y.c()), C.byref(a.c()), 99) == F.RCV_ERR_ARG
    # step smaller than a row
    bad = a.c()
    bad.step = 8
    assert F.lib.rcv_gaussian_blur(C.byref(bad), C.byref(R.Mat.new(8, 8, 3).c()), 5, 5, 0.0, 0.0) == F.RCV_ERR_SIZE
    # packed YUYV: the reference silently returns on a short src (videoio/mod.rs:346-348)
    # and would panic on a short dst; the ABI reports both
    src = np.zeros(6 * 4 * 2, np.uint8)
    dst = np.zeros(6 * 4 * 3, np.uint8)
    f = F.lib.rcv_yuyv_to_bgr_packed
    assert f(src.ctypes.data, src.size - 1, dst.ctypes.data, dst.size, 6, 4) == F.RCV_ERR_SIZE
    assert f(src.ctypes.data, src.size, dst.ctypes.data, dst.size - 1, 6, 4) == F.RCV_ERR_SIZE
    # Sobel wants single-channel f32
    assert F.lib.rcv_sobel_mag(C.byref(a.c()), C.byref(a.c()), None, None) == F.RCV_ERR_DEPTH
    # warpAffine: singular matrix
    f32 = R.Mat.new(8, 8, 1, R.F32)
    g32 = R.Mat.new(8, 8, 1, R.F32)
    m = (C.c_double * 6)(0, 0, 0, 0, 0, 0)
    assert F.lib.rcv_warp_affine(C.byref(f32.c()), C.byref(g32.c()), m, 0, 0.0) == F.RCV_ERR_ARG
    # in-place is rejected
    assert F.lib.rcv_gaussian_blur(C.byref(a.c()), C.byref(a.c()), 5, 5, 0.0, 0.0) == F.RCV_ERR_ARG


def test_host_side_matrix_helpers_match_oracle(oracle):
    from rustcv_b200 import imgproc

    M = imgproc.get_rotation_matrix_2d((41.0, 30.0), 15.0, 1.0)
    assert (M.ravel() == oracle.rotation_matrix(41.0, 30.0, 15.0)).all()
    from rustcv_b200 import _ffi as F

    im = (C.c_double * 6)()
    m = (C.c_double * 6)(*M.ravel())
    assert F.lib.rcv_invert_affine(m, im) == 0
    assert (np.array(im[:]) == oracle.invert_affine(M.ravel())).all()


def test_mat_mirrors_reference_semantics():
    # rustcv/src/core/mat.rs:18-51
    import rustcv_b200 as R

    m = R.Mat.new(4, 5, 3)
    assert (m.rows, m.cols, m.channels, m.step, m.data.size) == (4, 5, 3, 15, 60) and not m.is_empty()
    assert R.Mat.empty().is_empty()
    p = R.Mat.from_numpy_strided(np.arange(4 * 5 * 3, dtype=np.uint8).reshape(4, 5, 3), step=32)
    assert p.step == 32 and p.row_bytes(2).tolist() == list(range(30, 45))  # padding dropped
    assert (p.to_numpy() == np.arange(60, dtype=np.uint8).reshape(4, 5, 3)).all()
    # ensure_size: reallocate only when the byte length changes (rustcv-camera/src/mat.rs:65-74)
    buf = m.data
    m.ensure_size(5, 4, 3)
    assert m.data is buf and (m.rows, m.cols, m.step) == (5, 4, 12)
    m.ensure_size(6, 4, 3)
    assert m.data is not buf and m.data.size == 72


def test_fourcc_round_trip():
    # rustcv-camera/src/pixel_format.rs:148-172
    from rustcv_b200 import videoio

    assert videoio.YUYV == 0x56595559 and videoio.MJPEG == 0x47504A4D


def test_new_entry_points_validate_before_cuda_and_refuse_without_a_gpu():
    """Fused chains and the MJPEG branch: argument contracts are checked before any CUDA call (so they are testable
    here), and without an initialised B200 the calls refuse instead of falling back."""
    import rustcv_b200 as R
    from rustcv_b200 import _ffi as F

    yuyv = R.Mat.new(8, 8, 2)
    mag = R.Mat.new(8, 8, 1, R.F32)
    assert F.lib.rcv_yuyv_to_sobel_mag(C.byref(R.Mat.new(8, 9, 2).c()), C.byref(R.Mat.new(8, 9, 1, R.F32).c())) == F.RCV_ERR_SIZE
    assert F.lib.rcv_yuyv_to_sobel_mag(C.byref(R.Mat.new(8, 8, 3).c()), C.byref(mag.c())) == F.RCV_ERR_DEPTH
    assert F.lib.rcv_yuyv_to_sobel_mag(C.byref(yuyv.c()), C.byref(R.Mat.new(8, 8, 1).c())) == F.RCV_ERR_DEPTH
    assert F.lib.rcv_yuyv_to_sobel_mag(C.byref(yuyv.c()), C.byref(mag.c())) == F.RCV_ERR_NOT_INIT
    assert F.lib.rcv_yuyv_to_bgr_gaussian5(C.byref(yuyv.c()), C.byref(R.Mat.new(8, 8, 3).c())) == F.RCV_ERR_NOT_INIT
    assert F.lib.rcv_yuyv_to_sobel_mag_batch(None, None, 2) == F.RCV_ERR_ARG
    assert F.lib.rcv_yuyv_to_sobel_mag_batch(None, None, 0) == F.RCV_OK
    jpeg = np.frombuffer(b"\xff\xd8\xff\xe0" + bytes(60), np.uint8)
    w, h = C.c_int32(0), C.c_int32(0)
    assert F.lib.rcv_mjpeg_info(jpeg.ctypes.data, jpeg.size, C.byref(w), C.byref(h)) == F.RCV_ERR_NOT_INIT
    assert F.lib.rcv_mjpeg_info(None, 0, C.byref(w), C.byref(h)) == F.RCV_ERR_ARG
    assert F.lib.rcv_mjpeg_to_bgr(jpeg.ctypes.data, 2, C.byref(R.Mat.new(8, 8, 3).c())) == F.RCV_ERR_SIZE
    assert F.lib.rcv_mjpeg_to_bgr(jpeg.ctypes.data, jpeg.size, C.byref(R.Mat.new(8, 8, 1).c())) == F.RCV_ERR_DEPTH
    assert F.lib.rcv_mjpeg_to_bgr(jpeg.ctypes.data, jpeg.size, C.byref(R.Mat.new(8, 8, 3).c())) == F.RCV_ERR_NOT_INIT


def test_round2_entry_points_validate_before_cuda_and_refuse_without_a_gpu():
    """Multi-GPU batches, the coefficient broadcast, host registration: contracts checked before any CUDA call,
    no fallback without an initialised B200."""
    import rustcv_b200 as R
    from rustcv_b200 import _ffi as F
    from rustcv_b200.mat import MatBatch

    a = [R.Mat.new(8, 8, 3) for _ in range(3)]
    b = [R.Mat.new(8, 8, 3) for _ in range(3)]
    sa, sb = MatBatch.of(a), MatBatch.of(b)
    f = F.lib.rcv_gaussian_blur_batch_multi
    assert f(sa.arr, sb.arr, 3, 0, 5, 5, 0.0, 0.0) == F.RCV_ERR_NOT_INIT
    assert b"no CPU fallback" in F.lib.rcv_last_error()
    assert f(None, None, 3, 0, 5, 5, 0.0, 0.0) == F.RCV_ERR_ARG
    assert f(None, None, 0, 0, 5, 5, 0.0, 0.0) == F.RCV_OK
    # every element of a batch is checked: geometry ...
    odd = MatBatch.of([a[0], a[1], R.Mat.new(8, 9, 3)])
    assert f(odd.arr, sb.arr, 3, 0, 5, 5, 0.0, 0.0) == F.RCV_ERR_SIZE
    # ... and aliasing: dst[2] is src[0] (no op here runs in place; frames of a batch run concurrently)
    alias = MatBatch.of([b[0], b[1], a[0]])
    assert f(sa.arr, alias.arr, 3, 0, 5, 5, 0.0, 0.0) == F.RCV_ERR_ARG
    assert F.lib.rcv_gaussian_blur_batch(sa.arr, alias.arr, 3, 5, 5, 0.0, 0.0) == F.RCV_ERR_ARG
    # overlapping (not identical) buffers are in-place too
    big = np.zeros(8 * 24 + 24, np.uint8)
    s = R.Mat.new(8, 8, 3)
    s.data = big[:192]
    d = R.Mat.new(8, 8, 3)
    d.data = big[24:216]
    assert F.lib.rcv_cvt_color(C.byref(s.c()), C.byref(d.c()), F.COLOR_RGB2BGR) == F.RCV_ERR_ARG
    assert F.lib.rcv_gaussian_blur(C.byref(s.c()), C.byref(d.c()), 3, 3, 0.0, 0.0) == F.RCV_ERR_ARG
    # the fused YUYV -> GaussianBlur chain rejects odd widths (the blur would read an unconverted column)
    assert F.lib.rcv_yuyv_to_bgr_gaussian5(C.byref(R.Mat.new(8, 9, 2).c()), C.byref(R.Mat.new(8, 9, 3).c())) == F.RCV_ERR_SIZE
    # broadcast / registration / placement need an initialised GPU; argument errors come first
    co = (C.c_float * 4)(1, 2, 3, 4)
    assert F.lib.rcv_set_kernel_broadcast(None, 4, 0, 0, None) == F.RCV_ERR_ARG
    assert F.lib.rcv_set_kernel_broadcast(co, 0, 0, 0, None) == F.RCV_ERR_ARG
    assert F.lib.rcv_set_kernel_broadcast(co, 1000, 0, 0, None) == F.RCV_ERR_ARG
    assert F.lib.rcv_set_kernel_broadcast(co, 4, 0, 0, None) == F.RCV_ERR_NOT_INIT
    assert F.lib.rcv_host_register(None, 16) == F.RCV_ERR_ARG
    assert F.lib.rcv_host_register(big.ctypes.data, big.size) == F.RCV_ERR_NOT_INIT
    assert F.lib.rcv_host_unregister(big.ctypes.data) == F.RCV_OK  # nothing registered: nothing to undo
    p = C.c_void_p()
    assert F.lib.rcv_pinned_alloc_on(0, C.byref(p), 64) == F.RCV_ERR_NOT_INIT
    assert F.lib.rcv_pinned_free(big.ctypes.data) == F.RCV_ERR_ARG  # not a pointer of rcv_pinned_alloc
    taps = (C.c_int32 * 3)(64, 128, 64)
    g = F.lib.rcv_sep_filter2d_q8_batch_multi
    assert g(sa.arr, sb.arr, 3, 0, taps, 3, None, 3) == F.RCV_ERR_ARG  # kx without ky
    assert g(sa.arr, sb.arr, 3, 0, None, 0, None, 3) == F.RCV_ERR_ARG  # bank taps need their counts


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_init_multi_fails_loudly_without_a_gpu():
    from rustcv_b200 import _ffi as F

    assert F.lib.rcv_init_multi(0) == F.RCV_ERR_CUDA
    assert F.lib.rcv_init(-1) == F.RCV_ERR_CUDA  # RCV_DEVICE / GPU 0: still no GPU
