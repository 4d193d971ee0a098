// cvt.cu -- pixel-format conversion kernels (cvtColor family).
//
// These are the reference's real per-pixel loops:
//   YUYV->BGR  rustcv/src/videoio/mod.rs:344-371  (twin rustcv-camera/src/decode.rs:160-191)
//   clamp      rustcv/src/videoio/mod.rs:373-382
//   BGRA->BGR  rustcv/src/videoio/mod.rs:385-399  (twin decode.rs:200-207)
//   RGB<->BGR  rustcv-camera/src/decode.rs:213-219
//   BGR->XRGB  rustcv/src/highgui/mod.rs:125-141
//   NV12->BGR  rustcv-backend-msmf/examples/camera_view/convert.rs:46-86
// All HBM-bound byte shuffling: one 16-byte-vector kernel per format when base
// and step are 16-byte aligned, a scalar kernel otherwise.  Frames of a batch
// are blockIdx.z.
#include "rcv_internal.cuh"
#include "cvt_math.cuh"

namespace rcv {

// BT.601 integer formula, i32 math, arithmetic >> on negatives (videoio/mod.rs:352-369).
__device__ __forceinline__ uint32_t clamp_u8(int v) { return (uint32_t)min(max(v, 0), 255); }

struct Bgr2 {
  uint32_t b0, g0, r0, b1, g1, r1;
};

__device__ __forceinline__ Bgr2 yuv_pair(int y0, int ub, int y1, int vb) {
  int u = ub - 128, v = vb - 128;
  int c0 = 298 * (y0 - 16) + 128, c1 = 298 * (y1 - 16) + 128;
  int db = 516 * u, dg = -100 * u - 208 * v, dr = 409 * v;
  Bgr2 o;
  o.b0 = clamp_u8((c0 + db) >> 8);
  o.g0 = clamp_u8((c0 + dg) >> 8);
  o.r0 = clamp_u8((c0 + dr) >> 8);
  o.b1 = clamp_u8((c1 + db) >> 8);
  o.g1 = clamp_u8((c1 + dg) >> 8);
  o.r1 = clamp_u8((c1 + dr) >> 8);
  return o;
}

struct CvtArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *dst;
  size_t dstep, dfs;
  int rows, cols;
  int per_row;  // work items (threads) per row
  int mode;     // how threads map to (row, item): see cvt_index
  int rpb;      // mode 1: rows per block
  int magic;    // mode 1: ceil(2^16 / per_row)
};

// Thread -> (row, item).  Frames of a batch are blockIdx.y in every mode.
//   mode 1  narrow rows (per_row <= blockDim.x; a 640-wide YUYV row is only 40 vector items): a block takes
//           rpb = blockDim.x / per_row whole rows; row-in-block = (t * magic) >> 16, exact for t < 512
//   mode 2  wide rows: blockIdx.x walks the items of a row, blockIdx.z is the row (rows <= 65535)
//   mode 0  anything else: threads numbered over (row, item) with a 64-bit division
// (the division of mode 0 was a quarter of the YUYV->BGR kernel's instructions)
__device__ __forceinline__ bool cvt_index(const CvtArgs &a, int &r, int &i) {
  if (a.mode == 1) {
    const int t = (int)threadIdx.x, lr = (t * a.magic) >> 16;
    i = t - lr * a.per_row;
    r = (int)blockIdx.x * a.rpb + lr;
    return lr < a.rpb && r < a.rows;
  }
  if (a.mode == 2) {
    i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    r = (int)blockIdx.z;
    return i < a.per_row;
  }
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.rows * a.per_row) return false;
  r = (int)(t / a.per_row);
  i = (int)(t - (long long)r * a.per_row);
  return true;
}

// ---- scalar kernels: one thread per pixel / macro-pixel ------------------------------
template <int CODE>
__global__ void __launch_bounds__(256) k_cvt_scalar(CvtArgs a) {
  int r, i;
  if (!cvt_index(a, r, i)) return;
  const uint8_t *s = a.src + (size_t)blockIdx.y * a.sfs + (size_t)r * a.sstep;
  uint8_t *d = a.dst + (size_t)blockIdx.y * a.dfs + (size_t)r * a.dstep;
  if (CODE == RCV_COLOR_YUYV2BGR || CODE == RCV_COLOR_UYVY2BGR || CODE == RCV_COLOR_YUYV2GRAY) {
    if (i >= a.cols / 2) return;
    int q0 = s[i * 4], q1 = s[i * 4 + 1], q2 = s[i * 4 + 2], q3 = s[i * 4 + 3];
    Bgr2 o = (CODE == RCV_COLOR_UYVY2BGR) ? yuv_pair(q1, q0, q3, q2) : yuv_pair(q0, q1, q2, q3);
    if (CODE == RCV_COLOR_YUYV2GRAY) {
      d[i * 2] = (uint8_t)gray_of(o.b0, o.g0, o.r0);
      d[i * 2 + 1] = (uint8_t)gray_of(o.b1, o.g1, o.r1);
    } else {
      d[i * 6 + 0] = (uint8_t)o.b0;
      d[i * 6 + 1] = (uint8_t)o.g0;
      d[i * 6 + 2] = (uint8_t)o.r0;
      d[i * 6 + 3] = (uint8_t)o.b1;
      d[i * 6 + 4] = (uint8_t)o.g1;
      d[i * 6 + 5] = (uint8_t)o.r1;
    }
  } else {
    if (i >= a.cols) return;
    if (CODE == RCV_COLOR_BGRA2BGR) {
      d[i * 3] = s[i * 4];
      d[i * 3 + 1] = s[i * 4 + 1];
      d[i * 3 + 2] = s[i * 4 + 2];
    } else if (CODE == RCV_COLOR_RGB2BGR) {
      uint8_t b0 = s[i * 3], b1 = s[i * 3 + 1], b2 = s[i * 3 + 2];
      d[i * 3] = b2;
      d[i * 3 + 1] = b1;
      d[i * 3 + 2] = b0;
    } else if (CODE == RCV_COLOR_BGR2GRAY) {
      d[i] = (uint8_t)gray_of(s[i * 3], s[i * 3 + 1], s[i * 3 + 2]);
    } else if (CODE == RCV_COLOR_BGR2XRGB32) {
      ((uint32_t *)d)[i] = ((uint32_t)s[i * 3 + 2] << 16) | ((uint32_t)s[i * 3 + 1] << 8) | s[i * 3];
    }
  }
}

// ---- vector kernels ---------------------------------------------------------------------
// byte k of a little-endian word array
__device__ __forceinline__ uint32_t byte_of(const uint32_t *w, int k) { return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu; }


// YUYV/UYVY -> BGR: a thread converts 8 macro-pixels: 32 B in (2 x LDG.128) -> 48 B out (3 x STG.128).
// -> GRAY: 16 px -> 16 B out.
template <int CODE>
__global__ void __launch_bounds__(128) k_yuv422_vec(CvtArgs a) {
  int r, g;
  if (!cvt_index(a, r, g)) return;
  const uint8_t *s = a.src + (size_t)blockIdx.y * a.sfs + (size_t)r * a.sstep;
  uint8_t *d = a.dst + (size_t)blockIdx.y * a.dfs + (size_t)r * a.dstep;
  int pairs = a.cols / 2;
  if (g * 8 >= pairs) return;
  if (g * 8 + 8 <= pairs) {
    uint4 q0 = __ldg((const uint4 *)(s + (size_t)g * 32));
    uint4 q1 = __ldg((const uint4 *)(s + (size_t)g * 32 + 16));
    uint32_t in[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    if (CODE == RCV_COLOR_YUYV2GRAY) {
      uint32_t gr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // two macro-pixels -> 4 gray bytes, two pixels per operation (cvt_math.cuh)
        uint32_t b0, g0, r0, b1, g1, r1;
        yuv_word_pairs<false>(in[2 * k], b0, g0, r0);
        yuv_word_pairs<false>(in[2 * k + 1], b1, g1, r1);
        gr[k] = __byte_perm(gray_pair(b0, g0, r0), gray_pair(b1, g1, r1), 0x6420);
      }
      *(uint4 *)(d + (size_t)g * 16) = make_uint4(gr[0], gr[1], gr[2], gr[3]);
    } else {
      uint32_t out[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // two macro-pixels -> 12 bytes = 3 words
        constexpr bool U = CODE == RCV_COLOR_UYVY2BGR;
        uint32_t b0, g0, r0, b1, g1, r1;  // (pixel 0, pixel 1) of each channel as 16-bit lanes, values 0..255
        yuv_word_pairs<U>(in[2 * k], b0, g0, r0);
        yuv_word_pairs<U>(in[2 * k + 1], b1, g1, r1);
        const uint32_t X = __byte_perm(g0, r0, 0x6240);    // G0 R0 G1 R1
        const uint32_t Y = __byte_perm(b1, g1, 0x6240);    // B0' G0' B1' G1'
        out[3 * k] = __byte_perm(b0, X, 0x2540);           // B0 G0 R0 B1
        out[3 * k + 1] = __byte_perm(X, Y, 0x5432);        // G1 R1 B0' G0'
        out[3 * k + 2] = __byte_perm(r1, Y, 0x2760);       // R0' B1' G1' R1'
      }
      uint4 *dp = (uint4 *)(d + (size_t)g * 48);
      dp[0] = make_uint4(out[0], out[1], out[2], out[3]);
      dp[1] = make_uint4(out[4], out[5], out[6], out[7]);
      dp[2] = make_uint4(out[8], out[9], out[10], out[11]);
    }
  } else {
    for (int i = g * 8; i < pairs; ++i) {
      int q0 = s[i * 4], q1 = s[i * 4 + 1], q2 = s[i * 4 + 2], q3 = s[i * 4 + 3];
      Bgr2 o = (CODE == RCV_COLOR_UYVY2BGR) ? yuv_pair(q1, q0, q3, q2) : yuv_pair(q0, q1, q2, q3);
      if (CODE == RCV_COLOR_YUYV2GRAY) {
        d[i * 2] = (uint8_t)gray_of(o.b0, o.g0, o.r0);
        d[i * 2 + 1] = (uint8_t)gray_of(o.b1, o.g1, o.r1);
      } else {
        d[i * 6 + 0] = (uint8_t)o.b0;
        d[i * 6 + 1] = (uint8_t)o.g0;
        d[i * 6 + 2] = (uint8_t)o.r0;
        d[i * 6 + 3] = (uint8_t)o.b1;
        d[i * 6 + 4] = (uint8_t)o.g1;
        d[i * 6 + 5] = (uint8_t)o.r1;
      }
    }
  }
}

// 16 pixels per thread for the 3/4-byte formats.
//   BGRA2BGR   64 B in -> 48 B out      RGB2BGR 48 -> 48
//   BGR2GRAY   48 B in -> 16 B out      BGR2XRGB32 48 -> 64
template <int CODE>
__global__ void __launch_bounds__(128) k_px16_vec(CvtArgs a) {
  constexpr int IN_B = (CODE == RCV_COLOR_BGRA2BGR) ? 64 : 48;
  constexpr int OUT_B = (CODE == RCV_COLOR_BGRA2BGR || CODE == RCV_COLOR_RGB2BGR) ? 48
                        : (CODE == RCV_COLOR_BGR2GRAY)                            ? 16
                                                                                  : 64;
  constexpr int IN_PX = IN_B / 16, OUT_PX = OUT_B / 16;  // bytes per pixel
  int r, g;
  if (!cvt_index(a, r, g)) return;
  const uint8_t *s = a.src + (size_t)blockIdx.y * a.sfs + (size_t)r * a.sstep;
  uint8_t *d = a.dst + (size_t)blockIdx.y * a.dfs + (size_t)r * a.dstep;
  if (g * 16 >= a.cols) return;
  if (g * 16 + 16 <= a.cols) {
    uint32_t in[IN_B / 4], out[OUT_B / 4];
    const uint4 *sp = (const uint4 *)(s + (size_t)g * IN_B);
#pragma unroll
    for (int k = 0; k < IN_B / 16; ++k) {
      uint4 q = __ldg(sp + k);
      in[4 * k] = q.x;
      in[4 * k + 1] = q.y;
      in[4 * k + 2] = q.z;
      in[4 * k + 3] = q.w;
    }
    if (CODE == RCV_COLOR_BGR2XRGB32) {
#pragma unroll
      for (int p = 0; p < 16; ++p)
        out[p] = (byte_of(in, p * 3 + 2) << 16) | (byte_of(in, p * 3 + 1) << 8) | byte_of(in, p * 3);
    } else if (CODE == RCV_COLOR_BGR2GRAY) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int p = 4 * k + j;
          w |= gray_of(byte_of(in, p * 3), byte_of(in, p * 3 + 1), byte_of(in, p * 3 + 2)) << (8 * j);
        }
        out[k] = w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < OUT_B / 4; ++k) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int ob = 4 * k + j;  // output byte index
          int p = ob / 3, c = ob % 3;
          int sc = (CODE == RCV_COLOR_RGB2BGR) ? 2 - c : c;
          w |= byte_of(in, p * IN_PX + sc) << (8 * j);
        }
        out[k] = w;
      }
    }
    uint4 *dp = (uint4 *)(d + (size_t)g * OUT_B);
#pragma unroll
    for (int k = 0; k < OUT_B / 16; ++k) dp[k] = make_uint4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
  } else {
    for (int i = g * 16; i < a.cols; ++i) {
      const uint8_t *sp = s + (size_t)i * IN_PX;
      if (CODE == RCV_COLOR_BGRA2BGR) {
        d[i * 3] = sp[0];
        d[i * 3 + 1] = sp[1];
        d[i * 3 + 2] = sp[2];
      } else if (CODE == RCV_COLOR_RGB2BGR) {
        uint8_t b0 = sp[0], b1 = sp[1], b2 = sp[2];
        d[i * 3] = b2;
        d[i * 3 + 1] = b1;
        d[i * 3 + 2] = b0;
      } else if (CODE == RCV_COLOR_BGR2GRAY) {
        d[i] = (uint8_t)gray_of(sp[0], sp[1], sp[2]);
      } else {
        ((uint32_t *)d)[i] = ((uint32_t)sp[2] << 16) | ((uint32_t)sp[1] << 8) | sp[0];
      }
    }
  }
  (void)OUT_PX;
}

// NV12 -> BGR, per-pixel formula of convert.rs:46-86 (uv_row = row/2, uv_col = col/2).
struct Nv12Args {
  const uint8_t *y;
  size_t ystep;
  const uint8_t *uv;
  size_t uvstep;
  uint8_t *dst;
  size_t dstep;
  int rows, cols;
};

__global__ void __launch_bounds__(256) k_nv12(Nv12Args a) {
  int r = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // pixel pair
  int c0 = i * 2;
  if (c0 >= a.cols) return;
  const uint8_t *yr = a.y + (size_t)r * a.ystep;
  const uint8_t *uvr = a.uv + (size_t)(r / 2) * a.uvstep;
  uint8_t *d = a.dst + (size_t)r * a.dstep;
  int y0 = yr[c0];
  int y1 = (c0 + 1 < a.cols) ? yr[c0 + 1] : 16;
  Bgr2 o = yuv_pair(y0, uvr[c0], y1, uvr[c0 + 1]);
  d[c0 * 3] = (uint8_t)o.b0;
  d[c0 * 3 + 1] = (uint8_t)o.g0;
  d[c0 * 3 + 2] = (uint8_t)o.r0;
  if (c0 + 1 < a.cols) {
    d[c0 * 3 + 3] = (uint8_t)o.b1;
    d[c0 * 3 + 4] = (uint8_t)o.g1;
    d[c0 * 3 + 5] = (uint8_t)o.r1;
  }
}

static bool aligned16(const DBatch &b) {
  return (((uintptr_t)b.v.data | b.v.step | b.frame_stride) & 15) == 0;
}

template <int CODE>
static int launch_code(const DBatch &src, const DBatch &dst, cudaStream_t s) {
  CvtArgs a{src.v.data, src.v.step, src.frame_stride, dst.v.data, dst.v.step, dst.frame_stride, src.v.rows, src.v.cols, 0, 0, 0, 0};
  bool vec = aligned16(src) && aligned16(dst);
  constexpr bool yuv = (CODE == RCV_COLOR_YUYV2BGR || CODE == RCV_COLOR_UYVY2BGR || CODE == RCV_COLOR_YUYV2GRAY);
  const int threads = vec ? 128 : 256;
  if (vec)
    a.per_row = yuv ? ceil_div(src.v.cols / 2, 8) : ceil_div(src.v.cols, 16);
  else
    a.per_row = yuv ? src.v.cols / 2 : src.v.cols;
  if (a.per_row == 0) return RCV_OK;
  dim3 grid;
  if (a.per_row <= threads) {
    a.mode = 1;
    a.rpb = threads / a.per_row;
    a.magic = (65536 + a.per_row - 1) / a.per_row;
    grid = dim3((unsigned)ceil_div(a.rows, a.rpb), src.n, 1);
  } else if (a.rows <= 65535) {
    a.mode = 2;
    grid = dim3((unsigned)ceil_div(a.per_row, threads), src.n, (unsigned)a.rows);
  } else {
    a.mode = 0;
    const long long total = (long long)a.rows * a.per_row;
    const long long blocks = (total + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return fail(RCV_ERR_UNSUPPORTED, "image too large");
    grid = dim3((unsigned)blocks, src.n, 1);
  }
  if (vec) {
    if constexpr (yuv)
      k_yuv422_vec<CODE><<<grid, threads, 0, s>>>(a);
    else
      k_px16_vec<CODE><<<grid, threads, 0, s>>>(a);
  } else {
    k_cvt_scalar<CODE><<<grid, threads, 0, s>>>(a);
  }
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

int launch_cvt(Ctx *, const DBatch &src, const DBatch &dst, int code, cudaStream_t s) {
  if (src.n > 65535) return fail(RCV_ERR_UNSUPPORTED, "frames > 65535");
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  switch (code) {
    case RCV_COLOR_YUYV2BGR: return launch_code<RCV_COLOR_YUYV2BGR>(src, dst, s);
    case RCV_COLOR_UYVY2BGR: return launch_code<RCV_COLOR_UYVY2BGR>(src, dst, s);
    case RCV_COLOR_BGRA2BGR: return launch_code<RCV_COLOR_BGRA2BGR>(src, dst, s);
    case RCV_COLOR_RGB2BGR: return launch_code<RCV_COLOR_RGB2BGR>(src, dst, s);
    case RCV_COLOR_BGR2GRAY: return launch_code<RCV_COLOR_BGR2GRAY>(src, dst, s);
    case RCV_COLOR_BGR2XRGB32: return launch_code<RCV_COLOR_BGR2XRGB32>(src, dst, s);
    case RCV_COLOR_YUYV2GRAY: return launch_code<RCV_COLOR_YUYV2GRAY>(src, dst, s);
  }
  return fail(RCV_ERR_ARG, "unknown colour conversion code %d", code);
}

// ---- convertTo: u8 <-> f32 with scale/offset, 16 elements per thread ------------------------------
struct ConvArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *dst;
  size_t dstep, dfs;
  int rows, ncols;  // ncols = cols * channels
  int per_row;
  float a, b;
};

template <typename TS, typename TD>
__device__ __forceinline__ TD conv_one(TS v, float a, float b) {
  const float r = fmaf((float)v, a, b);
  if (sizeof(TD) == 1) return (TD)min(max(__float2int_rn(r), 0), 255);
  return (TD)r;
}

template <typename TS, typename TD, bool VEC>
__global__ void __launch_bounds__(256) k_convert(ConvArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.rows * a.per_row) return;
  const int r = (int)(t / a.per_row), g = (int)(t - (long long)r * a.per_row);
  const TS *s = (const TS *)(a.src + (size_t)blockIdx.y * a.sfs + (size_t)r * a.sstep);
  TD *d = (TD *)(a.dst + (size_t)blockIdx.y * a.dfs + (size_t)r * a.dstep);
  const int x0 = g * 16;
  if (VEC && x0 + 16 <= a.ncols) {
    __align__(16) TS in[16];
    __align__(16) TD out[16];
    const uint4 *sp = (const uint4 *)(s + x0);
#pragma unroll
    for (int k = 0; k < (int)sizeof(TS); ++k) ((uint4 *)in)[k] = __ldg(sp + k);
#pragma unroll
    for (int k = 0; k < 16; ++k) out[k] = conv_one<TS, TD>(in[k], a.a, a.b);
    uint4 *dp = (uint4 *)(d + x0);
#pragma unroll
    for (int k = 0; k < (int)sizeof(TD); ++k) dp[k] = ((const uint4 *)out)[k];
  } else {
    for (int x = x0; x < min(x0 + 16, a.ncols); ++x) d[x] = conv_one<TS, TD>(s[x], a.a, a.b);
  }
}

int launch_convert(Ctx *, const DBatch &src, const DBatch &dst, double alpha, double beta, cudaStream_t s) {
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  if (src.n > 65535) return fail(RCV_ERR_UNSUPPORTED, "frames > 65535");
  ConvArgs a{src.v.data, src.v.step, src.frame_stride, dst.v.data, dst.v.step, dst.frame_stride, src.v.rows,
             src.v.cols * src.v.cn, 0, (float)alpha, (float)beta};
  a.per_row = ceil_div(a.ncols, 16);
  const long long blocks = ((long long)a.rows * a.per_row + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(RCV_ERR_UNSUPPORTED, "image too large");
  dim3 grid((unsigned)blocks, src.n, 1);
  const bool vec = aligned16(src) && aligned16(dst);
  const bool sf = src.v.depth == RCV_F32, df = dst.v.depth == RCV_F32;
#define RCV_CONV(TS, TD)                                         \
  do {                                                           \
    if (vec)                                                     \
      k_convert<TS, TD, true><<<grid, 256, 0, s>>>(a);           \
    else                                                         \
      k_convert<TS, TD, false><<<grid, 256, 0, s>>>(a);          \
  } while (0)
  if (!sf && df)
    RCV_CONV(uint8_t, float);
  else if (sf && !df)
    RCV_CONV(float, uint8_t);
  else if (sf && df)
    RCV_CONV(float, float);
  else
    RCV_CONV(uint8_t, uint8_t);
#undef RCV_CONV
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

// NV12 vector kernel: a thread converts 16 pixels of one row -- 16 B of Y, the 16 B (8 U,V pairs) of the chroma
// row r/2 under them, 48 B out.  One PRMT interleaves (Y0, U, Y1, V) into the word layout of a YUYV macro-pixel,
// after which the arithmetic and the packing are k_yuv422_vec's.
__global__ void __launch_bounds__(128) k_nv12_vec(Nv12Args a, int per_row) {
  const int g = (int)(blockIdx.x * blockDim.x + threadIdx.x), r = (int)blockIdx.y;
  if (g >= per_row) return;
  const uint4 yq = __ldg((const uint4 *)(a.y + (size_t)r * a.ystep + (size_t)g * 16));
  const uint4 cq = __ldg((const uint4 *)(a.uv + (size_t)(r >> 1) * a.uvstep + (size_t)g * 16));
  const uint32_t yw[4] = {yq.x, yq.y, yq.z, yq.w}, cw[4] = {cq.x, cq.y, cq.z, cq.w};
  uint32_t out[12];
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // 4 pixels = 2 macro-pixels -> 12 bytes
    const uint32_t m0 = __byte_perm(yw[k], cw[k], 0x5140);  // Y0 U0 Y1 V0
    const uint32_t m1 = __byte_perm(yw[k], cw[k], 0x7362);  // Y2 U1 Y3 V1
    uint32_t b0, g0, r0, b1, g1, r1;
    yuv_word_pairs<false>(m0, b0, g0, r0);
    yuv_word_pairs<false>(m1, b1, g1, r1);
    const uint32_t X = __byte_perm(g0, r0, 0x6240);
    const uint32_t Y = __byte_perm(b1, g1, 0x6240);
    out[3 * k] = __byte_perm(b0, X, 0x2540);
    out[3 * k + 1] = __byte_perm(X, Y, 0x5432);
    out[3 * k + 2] = __byte_perm(r1, Y, 0x2760);
  }
  uint4 *dp = (uint4 *)(a.dst + (size_t)r * a.dstep + (size_t)g * 48);
  dp[0] = make_uint4(out[0], out[1], out[2], out[3]);
  dp[1] = make_uint4(out[4], out[5], out[6], out[7]);
  dp[2] = make_uint4(out[8], out[9], out[10], out[11]);
}

int launch_nv12(Ctx *, const DView &y, const DView &uv, const DView &dst, cudaStream_t s) {
  if (y.rows > 65535) return fail(RCV_ERR_UNSUPPORTED, "rows > 65535");
  if (y.rows == 0 || y.cols == 0) return RCV_OK;
  Nv12Args a{y.data, y.step, uv.data, uv.step, dst.data, dst.step, y.rows, y.cols};
  const bool al = ((((uintptr_t)y.data | y.step | (uintptr_t)uv.data | uv.step | (uintptr_t)dst.data | dst.step) & 15) == 0);
  if (al && (y.cols & 15) == 0) {
    const int per_row = y.cols / 16;
    dim3 grid(ceil_div(per_row, 128), y.rows, 1);
    k_nv12_vec<<<grid, 128, 0, s>>>(a, per_row);
  } else {
    dim3 grid(ceil_div(ceil_div(y.cols, 2), 256), y.rows, 1);
    k_nv12<<<grid, 256, 0, s>>>(a);
  }
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

}  // namespace rcv
