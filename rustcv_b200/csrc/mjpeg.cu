// mjpeg.cu -- the MJPEG branch of read() on the GPU.
//
// The reference decodes MJPG frames with TurboJPEG straight into the Mat, BGR, pitch = Mat.step
// (rustcv/src/videoio/mod.rs:205-232; twin rustcv-camera/src/decode.rs:93-121).  Here the same two steps --
// read the header, decompress to BGR at the Mat's pitch -- are split in two here:
//   1. entropy decoding + IDCT: nvJPEG (CUDA toolkit library, linked statically; NOT a kernel of this repo:
//      calling it counts like calling cuBLAS), asked for the planar Y / Cb / Cr samples at their native
//      (subsampled) resolution;
//   2. chroma upsampling + YCbCr -> BGR: k_ycc_to_bgr below, a restatement of what TurboJPEG does by default --
//      libjpeg-turbo's "fancy" triangle-filter upsampling (jdsample.c: h2v1_fancy_upsample, h2v2_fancy_upsample)
//      and its fixed-point colour conversion (jdcolor.c: ycc_rgb_convert).  nvJPEG's own BGR output replicates
//      chroma instead, which differs from the reference's decoder by tens of levels at chroma edges (measured:
//      mean 2.0 / max 78 on a 4:2:0 frame); with this kernel the remaining difference is the IDCT's last bit
//      (mean ~0.5, max 3-4 levels on every subsampling, tests/test_parity_gpu.py).
// The decoded BGR lands in HBM (device Mat: it stays there for the imgproc calls that follow; host Mat: one D2H
// copy), so only the compressed frame crosses PCIe.  4:4:4, 4:2:2, 4:2:0 and grayscale frames take this path;
// the rarer samplings (4:4:0, 4:1:1, 4:1:0) fall back to nvJPEG's own BGR output.
#include "rcv_internal.cuh"

#include <nvjpeg.h>

namespace rcv {

struct JpegCtx {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
};

// ---------------------------------------------------------------------------------------
// chroma upsampling + colour conversion, libjpeg-turbo semantics
// ---------------------------------------------------------------------------------------
struct YccArgs {
  const uint8_t *y, *cb, *cr;
  size_t ystep, cstep;
  uint8_t *dst;
  size_t dstep;
  int rows, cols;    // luma / output size
  int crows, ccols;  // chroma plane size
};

// jdcolor.c ycc_rgb_convert: FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 32768, arithmetic right shifts,
// range-limited to 0..255.
__device__ __forceinline__ void ycc_pixel(int y, int cb, int cr, uint8_t *o) {
  const int u = cb - 128, v = cr - 128;
  const int r = y + ((91881 * v + 32768) >> 16);
  const int b = y + ((116130 * u + 32768) >> 16);
  const int g = y + ((-22554 * u + 32768 - 46802 * v) >> 16);
  o[0] = (uint8_t)min(max(b, 0), 255);
  o[1] = (uint8_t)min(max(g, 0), 255);
  o[2] = (uint8_t)min(max(r, 0), 255);
}

// MODE 0: 4:4:4 (no upsampling), 1: 4:2:2 (h2v1 fancy), 2: 4:2:0 (h2v2 fancy), 3: grayscale.
// A thread makes the two output pixels 2i, 2i+1 of one row (they share chroma column i).  The edge columns /
// rows of libjpeg's routines are what the general formulas give with clamped neighbours:
//   h2v1  even = (3 c[i] + c[i-1] + 1) >> 2                 odd = (3 c[i] + c[i+1] + 2) >> 2
//   h2v2  s[j] = 3 near[j] + far[j] (far = the chroma row on the other side of this output row)
//         even = (3 s[i] + s[i-1] + 8) >> 4                 odd = (3 s[i] + s[i+1] + 7) >> 4
template <int MODE>
__global__ void __launch_bounds__(256) k_ycc_to_bgr(const YccArgs a) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x), r = (int)blockIdx.y;
  const int x0 = 2 * i;
  if (x0 >= a.cols) return;
  const uint8_t *yr = a.y + (size_t)r * a.ystep;
  const int y0 = yr[x0], y1 = x0 + 1 < a.cols ? yr[x0 + 1] : 0;
  int cb0, cb1, cr0, cr1;
  if (MODE == 3) {
    cb0 = cb1 = cr0 = cr1 = 128;
  } else if (MODE == 0) {
    const uint8_t *b = a.cb + (size_t)r * a.cstep, *c = a.cr + (size_t)r * a.cstep;
    const int x1 = min(x0 + 1, a.cols - 1);
    cb0 = b[x0];
    cb1 = b[x1];
    cr0 = c[x0];
    cr1 = c[x1];
  } else {
    const int im = max(i - 1, 0), ip = min(i + 1, a.ccols - 1);
    if (MODE == 1) {
      const uint8_t *b = a.cb + (size_t)r * a.cstep, *c = a.cr + (size_t)r * a.cstep;
      cb0 = (3 * b[i] + b[im] + 1) >> 2;
      cb1 = (3 * b[i] + b[ip] + 2) >> 2;
      cr0 = (3 * c[i] + c[im] + 1) >> 2;
      cr1 = (3 * c[i] + c[ip] + 2) >> 2;
    } else {
      const int rn = r >> 1;                                                    // nearer chroma row
      const int rf = (r & 1) ? min(rn + 1, a.crows - 1) : max(rn - 1, 0);       // the row on the other side
      const uint8_t *bn = a.cb + (size_t)rn * a.cstep, *bf = a.cb + (size_t)rf * a.cstep;
      const uint8_t *cn = a.cr + (size_t)rn * a.cstep, *cf = a.cr + (size_t)rf * a.cstep;
      const int sb = 3 * bn[i] + bf[i], sbm = 3 * bn[im] + bf[im], sbp = 3 * bn[ip] + bf[ip];
      const int sc = 3 * cn[i] + cf[i], scm = 3 * cn[im] + cf[im], scp = 3 * cn[ip] + cf[ip];
      cb0 = (3 * sb + sbm + 8) >> 4;
      cb1 = (3 * sb + sbp + 7) >> 4;
      cr0 = (3 * sc + scm + 8) >> 4;
      cr1 = (3 * sc + scp + 7) >> 4;
    }
  }
  uint8_t *d = a.dst + (size_t)r * a.dstep + (size_t)x0 * 3;
  uint8_t px[6];
  ycc_pixel(y0, cb0, cr0, px);
  ycc_pixel(y1, cb1, cr1, px + 3);
  const int n = x0 + 1 < a.cols ? 6 : 3;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (k < n) d[k] = px[k];
}

static const char *nvjpeg_name(nvjpegStatus_t st) {
  switch (st) {
    case NVJPEG_STATUS_NOT_INITIALIZED: return "NOT_INITIALIZED";
    case NVJPEG_STATUS_INVALID_PARAMETER: return "INVALID_PARAMETER";
    case NVJPEG_STATUS_BAD_JPEG: return "BAD_JPEG";
    case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "JPEG_NOT_SUPPORTED";
    case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "ALLOCATOR_FAILURE";
    case NVJPEG_STATUS_EXECUTION_FAILED: return "EXECUTION_FAILED";
    case NVJPEG_STATUS_ARCH_MISMATCH: return "ARCH_MISMATCH";
    case NVJPEG_STATUS_INTERNAL_ERROR: return "INTERNAL_ERROR";
    case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "IMPLEMENTATION_NOT_SUPPORTED";
    default: return "error";
  }
}

static int nvjpeg_fail(nvjpegStatus_t st, const char *what) {
  const int code = (st == NVJPEG_STATUS_BAD_JPEG || st == NVJPEG_STATUS_INVALID_PARAMETER) ? RCV_ERR_ARG
                   : (st == NVJPEG_STATUS_JPEG_NOT_SUPPORTED || st == NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED)
                       ? RCV_ERR_UNSUPPORTED
                       : RCV_ERR_CUDA;
  return fail(code, "nvJPEG %s: %s", what, nvjpeg_name(st));
}

// the context's decoder, created on first use (callers hold c->mu)
static int jpeg_get(Ctx *c, JpegCtx **out) {
  if (!c->jpeg) {
    JpegCtx *j = new JpegCtx;
    nvjpegStatus_t st = nvjpegCreateSimple(&j->handle);
    if (st == NVJPEG_STATUS_SUCCESS) st = nvjpegJpegStateCreate(j->handle, &j->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
      if (j->handle) nvjpegDestroy(j->handle);
      delete j;
      return nvjpeg_fail(st, "init");
    }
    c->jpeg = j;
  }
  *out = (JpegCtx *)c->jpeg;
  return RCV_OK;
}

void jpeg_destroy(Ctx *c) {
  JpegCtx *j = (JpegCtx *)c->jpeg;
  if (!j) return;
  if (j->state) nvjpegJpegStateDestroy(j->state);
  if (j->handle) nvjpegDestroy(j->handle);
  delete j;
  c->jpeg = nullptr;
}

struct JpegInfo {
  int ncomp, w[NVJPEG_MAX_COMPONENT], h[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t ss;
};

static int jpeg_header(JpegCtx *j, const uint8_t *jpeg, size_t len, JpegInfo *o) {
  for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) o->w[k] = o->h[k] = 0;
  nvjpegStatus_t st = nvjpegGetImageInfo(j->handle, jpeg, len, &o->ncomp, &o->ss, o->w, o->h);
  if (st != NVJPEG_STATUS_SUCCESS) return nvjpeg_fail(st, "header");
  return RCV_OK;
}

int mjpeg_info(Ctx *c, const uint8_t *jpeg, size_t len, int *width, int *height) {
  JpegCtx *j = nullptr;
  RCV_TRY(jpeg_get(c, &j));
  JpegInfo in;
  RCV_TRY(jpeg_header(j, jpeg, len, &in));
  *width = in.w[0];
  *height = in.h[0];
  return RCV_OK;
}

// dst: device view, u8 C3, rows x cols equal to the JPEG's; BGR interleaved at dst.step.
int launch_mjpeg(Ctx *c, const uint8_t *jpeg, size_t len, const DView &dst, cudaStream_t s) {
  JpegCtx *j = nullptr;
  RCV_TRY(jpeg_get(c, &j));
  JpegInfo in;
  RCV_TRY(jpeg_header(j, jpeg, len, &in));
  nvjpegImage_t img;
  for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) {
    img.channel[k] = nullptr;
    img.pitch[k] = 0;
  }
  int mode = -1;
  if (in.ncomp == 1 || in.ss == NVJPEG_CSS_GRAY) mode = 3;
  else if (in.ncomp == 3 && in.ss == NVJPEG_CSS_444) mode = 0;
  else if (in.ncomp == 3 && in.ss == NVJPEG_CSS_422) mode = 1;
  else if (in.ncomp == 3 && in.ss == NVJPEG_CSS_420) mode = 2;
  if (opt_get("mjpeg.library_color", 0) != 0) mode = -1;
  if (mode < 0) {  // rarer samplings: the library's own upsampling + colour conversion
    img.channel[0] = dst.data;
    img.pitch[0] = dst.step;
    nvjpegStatus_t st = nvjpegDecode(j->handle, j->state, jpeg, len, NVJPEG_OUTPUT_BGRI, &img, s);
    if (st != NVJPEG_STATUS_SUCCESS) return nvjpeg_fail(st, "decode");
    count_launch();
    return RCV_OK;
  }
  // planar samples at native resolution into context scratch, then the upsampling + conversion kernel
  const int planes = mode == 3 ? 1 : 3;
  void *buf[3] = {nullptr, nullptr, nullptr};
  size_t pitch[3] = {0, 0, 0};
  for (int k = 0; k < planes; ++k) {
    pitch[k] = ((size_t)in.w[k] + 255) / 256 * 256;
    RCV_TRY(ctx_scratch(c, SCR_JPEG_Y + k, pitch[k] * (size_t)in.h[k], &buf[k]));
    img.channel[k] = (unsigned char *)buf[k];
    img.pitch[k] = pitch[k];
  }
  nvjpegStatus_t st = nvjpegDecode(j->handle, j->state, jpeg, len, mode == 3 ? NVJPEG_OUTPUT_Y : NVJPEG_OUTPUT_YUV, &img, s);
  if (st != NVJPEG_STATUS_SUCCESS) return nvjpeg_fail(st, "decode");
  count_launch();  // the library's device kernels (IDCT)
  YccArgs a;
  a.y = (const uint8_t *)buf[0];
  a.cb = (const uint8_t *)buf[1];
  a.cr = (const uint8_t *)buf[2];
  a.ystep = pitch[0];
  a.cstep = pitch[1];
  a.dst = dst.data;
  a.dstep = dst.step;
  a.rows = dst.rows;
  a.cols = dst.cols;
  a.crows = in.h[1];
  a.ccols = in.w[1];
  if (mode != 3 && pitch[2] != pitch[1]) return fail(RCV_ERR_UNSUPPORTED, "Cb and Cr planes differ in size");
  if (dst.rows > 65535) return fail(RCV_ERR_UNSUPPORTED, "rows > 65535");
  dim3 grid(ceil_div(ceil_div(dst.cols, 2), 256), dst.rows, 1);
  switch (mode) {
    case 0: k_ycc_to_bgr<0><<<grid, 256, 0, s>>>(a); break;
    case 1: k_ycc_to_bgr<1><<<grid, 256, 0, s>>>(a); break;
    case 2: k_ycc_to_bgr<2><<<grid, 256, 0, s>>>(a); break;
    default: k_ycc_to_bgr<3><<<grid, 256, 0, s>>>(a); break;
  }
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

}  // namespace rcv
