// filter.cu -- generic (any taps, any geometry) filters: separable u8 Q8, separable f32,
// dense filter2D, and the small-image Sobel.  These are the general-case companions of
// the strip-pipeline kernels in stencil.cu; they take every size/alignment, including
// 1-row / 1-column images, and define BORDER_REFLECT_101 by index arithmetic.
//
// The reference has no filters (rustcv/src/imgproc/mod.rs:1-4); semantics and operation
// order are the oracle's: orc_sepfilter_u8_q8, orc_sepfilter_f32, orc_filter2d_{f32,u8},
// orc_sobel3_f32 in oracle/rcv_oracle.c.
#include "rcv_internal.cuh"

#include <cstring>

#include <cmath>

namespace rcv {

constexpr int kMaxTaps = 31;

// ---------------------------------------------------------------------------------------
// separable filter: CTA tile of TBX element-columns x TBY rows.
//   stage 1  raw tile (+ halo, reflected) -> shared
//   stage 2  horizontal pass              -> shared (accumulator type)
//   stage 3  vertical pass                -> global
// "element column" = index into the row in elements (pixel*cn + channel); the horizontal
// tap j reads element column x + (j - rx)*cn.
// ---------------------------------------------------------------------------------------
template <typename T>
struct SepTraits;
template <>
struct SepTraits<uint8_t> {
  typedef uint32_t Acc;
  typedef int32_t Tap;
};
template <>
struct SepTraits<float> {
  typedef float Acc;
  typedef float Tap;
};

template <typename T>
struct SepArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *dst;
  size_t dstep, dfs;
  int rows, cols, cn;
  int kw, kh;
  typename SepTraits<T>::Tap kx[kMaxTaps + 1], ky[kMaxTaps + 1];
};

// ---------------------------------------------------------------------------------------
// register-tiled tap chains
// ---------------------------------------------------------------------------------------
// acc[m] (m < N) <- mac(k[j], v(m + j), acc[m]) for j = 0 .. kn-1 in ASCENDING j (the oracle's fmaf order), where
// v(i) = load(i), i = 0 .. kn + N - 2.  Every v(i) and every tap is loaded ONCE and used for up to N outputs, so
// the loop runs at ~(1 + 2/N) instructions per multiply-add instead of the 3 of a one-output-per-thread loop
// (load value, load tap, fma) -- the old general-case kernels were shared-memory-instruction bound at 3-5 % of the
// HBM roofline.  The last N taps live in a register ring indexed by compile-time slots (tap j in slot j % N).
// KN > 0: the tap count is a compile-time constant (everything unrolls, no predicates); KN == 0: run time.
template <int N, int KN, class Acc, class Tap, class Load, class Mac>
__device__ __forceinline__ void tap_chain(Acc (&acc)[N], const Tap *__restrict__ k, int kn_rt, Load load, Mac mac) {
  const int kn = KN > 0 ? KN : kn_rt;
  Tap kk[N];
#pragma unroll
  for (int m = 0; m < N; ++m) kk[m] = Tap(0);
  if constexpr (KN > 0) {
#pragma unroll
    for (int i = 0; i < KN + N - 1; ++i) {
      const Acc v = load(i);
      if (i < KN) kk[i % N] = k[i];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const int j = i - m;
        if (j >= 0 && j < KN) acc[m] = mac(kk[(j + N) % N], v, acc[m]);
      }
    }
  } else {
  // ramp-up: i = 0 .. N-2, outputs m <= i
#pragma unroll
  for (int i = 0; i < N - 1; ++i) {
    if (i < kn + N - 1) {
      const Acc v = load(i);
      if (i < kn) kk[i % N] = k[i];
#pragma unroll
      for (int m = 0; m <= i; ++m)
        if (i - m < kn) acc[m] = mac(kk[(i - m) % N], v, acc[m]);
    }
  }
  // steady state: N values per trip, every output takes every value (i0 = N-1 mod N throughout)
  int i0 = N - 1;
  for (; i0 + N - 1 <= kn - 1; i0 += N) {
#pragma unroll
    for (int u = 0; u < N; ++u) {
      const Acc v = load(i0 + u);
      kk[(N - 1 + u) % N] = k[i0 + u];
#pragma unroll
      for (int m = 0; m < N; ++m) acc[m] = mac(kk[(2 * N - 1 + u - m) % N], v, acc[m]);
    }
  }
  // ramp-down: the remaining i0 .. kn + N - 2 (at most 2N - 2 values)
#pragma unroll
  for (int u = 0; u < 2 * N - 2; ++u) {
    const int i = i0 + u;
    if (i < kn + N - 1) {
      const Acc v = load(i);
      if (i < kn) kk[(N - 1 + u) % N] = k[i];
#pragma unroll
      for (int m = 0; m < N; ++m)
        if (i - m < kn) acc[m] = mac(kk[(4 * N - 1 + u - m) % N], v, acc[m]);
    }
  }
  }
}

// lanes of a warp that work on a tile row: a lane owns one channel of a block of N pixels, so only whole pixels
constexpr int tile_lanes(int cn) { return 32 / cn * cn; }

// Outputs per thread: 8 rows in the vertical pass, 8 pixels in the horizontal pass.  (9 pixels would spread a
// warp's lanes over distinct shared-memory banks -- see kF2dN -- but the wider tile costs a resident CTA per SM here:
// measured slower for u8 and for 3-channel f32, profiles/r2_generic_kernels.txt.)
constexpr int kSepTBY = 32, kSepThreads = 256, kSepN = 8, kSepNH = 8;
constexpr int kSepMaxTBX = 32 * kSepNH;                    // tile width in element columns (270 for 3 channels)
constexpr int kSepMaxRW = kSepMaxTBX + (kMaxTaps - 1) * 4;  // widest raw tile row (elements)

// Stage 2 (horizontal): warp w takes tile rows w, w+8, ...; a lane computes the kSepNH outputs of ONE channel of a
// block of kSepNH pixels (element columns x0 + m*cn) from the kw + kSepNH - 1 raw values x0 + i*cn.
// Stage 3 (vertical): a thread computes kSepN consecutive rows of one element column from kh + kSepN - 1 values of
// the horizontally filtered tile.
template <typename T, int KN>
__global__ void __launch_bounds__(kSepThreads) k_sepfilter(const __grid_constant__ SepArgs<T> a) {
  typedef typename SepTraits<T>::Acc Acc;
  typedef typename SepTraits<T>::Tap Tap;
  extern __shared__ __align__(16) uint8_t smem[];
  const int kw = KN > 0 ? KN : a.kw, kh = KN > 0 ? KN : a.kh;
  const int rx = kw / 2, ry = kh / 2;
  const int lanes = 32 / a.cn * a.cn;
  const int TBX = lanes * kSepNH;          // tile width (element columns)
  const int RW = TBX + (kw - 1) * a.cn;    // raw tile width (elements)
  const int RH = kSepTBY + kh - 1;         // raw tile height
  T *raw = (T *)smem;
  Acc *mid = (Acc *)(smem + (((size_t)RW * RH * sizeof(T) + 15) & ~(size_t)15));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NWARP = kSepThreads / 32;

  const int ncols = a.cols * a.cn;
  const int ex0 = blockIdx.x * TBX;  // first element column of the tile
  const int y0 = blockIdx.y * kSepTBY;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  uint8_t *dst = a.dst + (size_t)blockIdx.z * a.dfs;

  // stage 1: raw[r][x] = src[reflect(y0 + r - ry)][reflect column of x]; the reflected source offset of every
  // raw-tile column is computed once per thread (no divisions in the row loop)
  constexpr int NX = (kSepMaxRW + 31) / 32;
  const bool interior_x = ex0 - rx * a.cn >= 0 && ex0 - rx * a.cn + RW <= ncols;
  int xoff[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    const int x = lane + 32 * i;
    xoff[i] = 0;
    if (x < RW) {
      const int ex = ex0 + x - rx * a.cn;
      if (interior_x) {  // the tile's columns and halo lie inside the row: no reflection, no division
        xoff[i] = ex;
      } else {
        const int px = ex >= 0 ? ex / a.cn : -((-ex + a.cn - 1) / a.cn);  // floor division
        xoff[i] = reflect101(px, a.cols) * a.cn + (ex - px * a.cn);
      }
    }
  }
  for (int r = warp; r < RH; r += NWARP) {
    const T *srow = (const T *)(src + (size_t)reflect101(y0 + r - ry, a.rows) * a.sstep);
    T *rrow = raw + r * RW;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int x = lane + 32 * i;
      if (x < RW) rrow[x] = srow[xoff[i]];
    }
  }
  __syncthreads();
  // stage 2: horizontal
  if (lane < lanes) {
    const int x0 = lane / a.cn * (kSepNH * a.cn) + lane % a.cn;
    for (int r = warp; r < RH; r += NWARP) {
      const T *p = raw + r * RW + x0;
      Acc acc[kSepNH];
#pragma unroll
      for (int m = 0; m < kSepNH; ++m) acc[m] = 0;
      const int cn = a.cn;
      if (sizeof(T) == 1)
        tap_chain<kSepNH, KN>(acc, (const Tap *)a.kx, kw, [&](int i) { return (Acc)p[i * cn]; },
                              [](Tap k, Acc v, Acc c) { return (Acc)(c + (Acc)k * v); });
      else
        tap_chain<kSepNH, KN>(acc, (const Tap *)a.kx, kw, [&](int i) { return (Acc)p[i * cn]; },
                              [](Tap k, Acc v, Acc c) { return (Acc)fmaf((float)k, (float)v, (float)c); });
      Acc *o = mid + r * TBX + x0;
#pragma unroll
      for (int m = 0; m < kSepNH; ++m) o[m * cn] = acc[m];
    }
  }
  __syncthreads();
  // stage 3: vertical -- thread t: element column t % TBX (+ multiples of the thread count), row group
  for (int task = threadIdx.x; task < TBX * (kSepTBY / kSepN); task += kSepThreads) {
    const int x = task % TBX, r0 = task / TBX * kSepN;
    const int ex = ex0 + x;
    if (ex >= ncols || y0 + r0 >= a.rows) continue;
    const Acc *p = mid + r0 * TBX + x;
    Acc acc[kSepN];
#pragma unroll
    for (int m = 0; m < kSepN; ++m) acc[m] = sizeof(T) == 1 ? (Acc)32768u : (Acc)0;
    if (sizeof(T) == 1)
      tap_chain<kSepN, KN>(acc, (const Tap *)a.ky, kh, [&](int i) { return p[i * TBX]; },
                           [](Tap k, Acc v, Acc c) { return (Acc)(c + (Acc)k * v); });
    else
      tap_chain<kSepN, KN>(acc, (const Tap *)a.ky, kh, [&](int i) { return p[i * TBX]; },
                           [](Tap k, Acc v, Acc c) { return (Acc)fmaf((float)k, (float)v, (float)c); });
#pragma unroll
    for (int m = 0; m < kSepN; ++m) {
      const int y = y0 + r0 + m;
      if (y >= a.rows) break;
      if (sizeof(T) == 1) {
        const uint32_t v = (uint32_t)acc[m] >> 16;
        ((uint8_t *)(dst + (size_t)y * a.dstep))[ex] = (uint8_t)(v > 255u ? 255u : v);
      } else {
        ((float *)(dst + (size_t)y * a.dstep))[ex] = (float)acc[m];
      }
    }
  }
}

template <typename T, int KN>
static int launch_sep_kn(const SepArgs<T> &a, int n, cudaStream_t s) {
  const int kw = KN > 0 ? KN : a.kw, kh = KN > 0 ? KN : a.kh;
  const int TBX = tile_lanes(a.cn) * kSepNH;
  const int RW = TBX + (kw - 1) * a.cn, RH = kSepTBY + kh - 1;
  const size_t smem = (((size_t)RW * RH * sizeof(T) + 15) & ~(size_t)15) + (size_t)TBX * RH * 4;
  auto kern = k_sepfilter<T, KN>;
  static size_t attr_smem[16] = {};  // per instantiation, per device: the largest size set so far
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_smem[dev & 15] < smem) {
    RCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[dev & 15] = smem;
  }
  const int ncols = a.cols * a.cn;
  dim3 grid(ceil_div(ncols, TBX), ceil_div(a.rows, kSepTBY), n);
  if (grid.y > 65535 || grid.z > 65535) return fail(RCV_ERR_UNSUPPORTED, "image too tall / batch too large");
  kern<<<grid, kSepThreads, smem, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

template <typename T>
static int launch_sep(const DBatch &src, const DBatch &dst, const typename SepTraits<T>::Tap *kx, int kw,
                      const typename SepTraits<T>::Tap *ky, int kh, cudaStream_t s) {
  if (kw < 1 || kh < 1 || kw > kMaxTaps || kh > kMaxTaps)
    return fail(RCV_ERR_ARG, "kernel size %dx%d outside 1..%d", kw, kh, kMaxTaps);
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  SepArgs<T> a;
  a.src = src.v.data;
  a.sstep = src.v.step;
  a.sfs = src.frame_stride;
  a.dst = dst.v.data;
  a.dstep = dst.v.step;
  a.dfs = dst.frame_stride;
  a.rows = src.v.rows;
  a.cols = src.v.cols;
  a.cn = src.v.cn;
  a.kw = kw;
  a.kh = kh;
  for (int i = 0; i < kw; ++i) a.kx[i] = kx[i];
  for (int i = 0; i < kh; ++i) a.ky[i] = ky[i];
  // square kernels of the common sizes: compile-time tap counts (fully unrolled chains, no predicates)
  if (kw == kh) {
    switch (kw) {
      case 3: return launch_sep_kn<T, 3>(a, src.n, s);
      case 5: return launch_sep_kn<T, 5>(a, src.n, s);
      case 7: return launch_sep_kn<T, 7>(a, src.n, s);
      case 9: return launch_sep_kn<T, 9>(a, src.n, s);
      case 11: return launch_sep_kn<T, 11>(a, src.n, s);
      case 13: return launch_sep_kn<T, 13>(a, src.n, s);
      case 15: return launch_sep_kn<T, 15>(a, src.n, s);
    }
  }
  return launch_sep_kn<T, 0>(a, src.n, s);
}

int launch_gauss5_strip(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s);
int launch_gauss3_strip(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s);
int launch_gaussq8_strip(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, const int32_t *ky, int ks,
                         cudaStream_t s);

// Symmetric non-negative taps summing to 256: the vertical sums fit 16-bit lanes, which is what the strip ops
// (and nothing else about a Gaussian) rely on.
static bool strip_taps_ok(const int32_t *k, int n) {
  int sum = 0;
  for (int i = 0; i < n; ++i) {
    if (k[i] < 0 || k[i] != k[n - 1 - i]) return false;
    sum += k[i];
  }
  return sum == 256;
}

// u8 separable filter with Q8 taps: the TMA strip ops when the taps allow it (5x5 / 3x3 binomial have their own
// ops, any other symmetric 3 / 5 / 7 taps GaussQ8Op), else the general kernel.
int launch_sepfilter_q8(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, int kw, const int32_t *ky,
                        int kh, cudaStream_t s) {
  if (kw == kh && (kw & 1) && kw >= 3 && kw <= 15 && opt_get("gauss.force_generic", 0) == 0 && strip_taps_ok(kx, kw) &&
      strip_taps_ok(ky, kh) && src.v.rows > 0 && src.v.cols > 0 && src.n > 0) {
    const bool binomial5 = kw == 5 && kx[0] == 16 && kx[1] == 64 && kx[2] == 96 && ky[0] == 16 && ky[1] == 64 && ky[2] == 96;
    if (binomial5) {
      int rc = launch_gauss5_strip(c, src, dst, s);
      if (rc != RCV_ERR_UNSUPPORTED) return rc;
    }
    const bool binomial3 = kw == 3 && kx[0] == 64 && kx[1] == 128 && ky[0] == 64 && ky[1] == 128;
    if (binomial3 && opt_get("gauss.no_binomial3", 0) == 0) {
      int rc = launch_gauss3_strip(c, src, dst, s);
      if (rc != RCV_ERR_UNSUPPORTED) return rc;
    }
    int rc = launch_gaussq8_strip(c, src, dst, kx, ky, kw, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }
  if (dst.windowed()) return RCV_ERR_UNSUPPORTED;  // whole-image kernel: the caller retries without a row window
  return launch_sep<uint8_t>(src, dst, kx, kw, ky, kh, s);
}

int launch_sepf32_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky, int kh,
                        cudaStream_t s);
int launch_filter2d_f32_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                              cudaStream_t s);
int launch_filter2d_u8_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                             cudaStream_t s);

int launch_sepfilter_f32(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky,
                         int kh, cudaStream_t s) {
  if (opt_get("sepf32.force_generic", 0) == 0 && src.v.rows > 0 && src.v.cols > 0 && src.n > 0) {
    int rc = launch_sepf32_strip(c, src, dst, kx, kw, ky, kh, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }
  if (dst.windowed()) return RCV_ERR_UNSUPPORTED;  // whole-image kernel: the caller retries without a row window
  return launch_sep<float>(src, dst, kx, kw, ky, kh, s);
}

// ---------------------------------------------------------------------------------------
// GaussianBlur front end (cv::GaussianBlur model, oracle: orc_gaussian_blur_{u8,f32})
// ---------------------------------------------------------------------------------------
int gaussian_ksize(double sigma, bool is_u8) {
  int k = (int)lrint(sigma * (is_u8 ? 3 : 4) * 2 + 1);
  return k | 1;
}

void gaussian_kernel_f64(int n, double sigma, double *kd) {
  static const double t1[] = {1.0};
  static const double t3[] = {0.25, 0.5, 0.25};
  static const double t5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
  static const double t7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
  if (sigma <= 0 && n <= 7 && (n & 1)) {
    const double *t = n == 1 ? t1 : n == 3 ? t3 : n == 5 ? t5 : t7;
    for (int i = 0; i < n; ++i) kd[i] = t[i];
    return;
  }
  double sig = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
  double scale2x = -0.5 / (sig * sig);
  double sum = 0;
  for (int i = 0; i < n; ++i) {
    double x = i - (n - 1) * 0.5;
    kd[i] = exp(scale2x * x * x);
    sum += kd[i];
  }
  for (int i = 0; i < n; ++i) kd[i] /= sum;
}

void gaussian_kernel_q8(int n, double sigma, int32_t *kq) {
  double kd[kMaxTaps + 1];
  gaussian_kernel_f64(n, sigma, kd);
  int n2 = n / 2;
  double err = 0;
  int sum = 0;
  for (int i = 0; i < n2; ++i) {
    double adj = kd[i] * 256.0 + err;
    int v0 = (int)lrint(adj);
    err = adj - v0;
    kq[i] = v0;
    kq[n - 1 - i] = v0;
    sum += v0;
  }
  kq[n2] = 256 - 2 * sum;
}

int launch_gaussian(Ctx *c, const DBatch &src, const DBatch &dst, int kw, int kh, double sx, double sy,
                    cudaStream_t s) {
  const bool u8 = src.v.depth == RCV_U8;
  if (sy <= 0) sy = sx;
  if (kw <= 0 && sx > 0) kw = gaussian_ksize(sx, u8);
  if (kh <= 0 && sy > 0) kh = gaussian_ksize(sy, u8);
  if (kw < 1 || kh < 1 || !(kw & 1) || !(kh & 1) || kw > kMaxTaps || kh > kMaxTaps)
    return fail(RCV_ERR_ARG, "GaussianBlur needs odd kernel sizes in 1..%d (got %dx%d)", kMaxTaps, kw, kh);
  if (u8) {
    int32_t kx[kMaxTaps + 1], ky[kMaxTaps + 1];
    gaussian_kernel_q8(kw, sx, kx);
    gaussian_kernel_q8(kh, sy, ky);
    return launch_sepfilter_q8(c, src, dst, kx, kw, ky, kh, s);
  }
  double kd[kMaxTaps + 1];
  float kx[kMaxTaps + 1], ky[kMaxTaps + 1];
  gaussian_kernel_f64(kw, sx, kd);
  for (int i = 0; i < kw; ++i) kx[i] = (float)kd[i];
  gaussian_kernel_f64(kh, sy, kd);
  for (int i = 0; i < kh; ++i) ky[i] = (float)kd[i];
  return launch_sepfilter_f32(c, src, dst, kx, kw, ky, kh, s);
}

// YUYV -> BGR -> GaussianBlur 5x5.  One fused strip kernel
// (strip_yuyv_gauss5.cu) whenever the strip path applies; otherwise (tiny, odd-width or unaligned images) the
// two stand-alone kernels over a device-resident intermediate.
int launch_yuyv_gauss5_strip(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s);

int launch_yuyv_gauss5(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s) {
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  if (opt_get("yuyvgauss.force_chain", 0) == 0) {
    int rc = launch_yuyv_gauss5_strip(c, src, dst, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }
  if (dst.windowed()) return RCV_ERR_UNSUPPORTED;
  DBatch tmp = dst;
  size_t pitch = (dst.v.row_bytes() + 255) / 256 * 256;
  size_t frame = pitch * (size_t)dst.v.rows;
  void *p = nullptr;
  RCV_TRY(ctx_scratch(c, SCR_FUSE_TMP, frame * src.n, &p));
  tmp.v.data = (uint8_t *)p;
  tmp.v.step = pitch;
  tmp.frame_stride = src.n > 1 ? frame : 0;
  RCV_TRY(launch_cvt(c, src, tmp, RCV_COLOR_YUYV2BGR, s));
  // an odd trailing column is left untouched by the conversion (cols/2 macro-pixels per row)
  return launch_gaussian(c, tmp, dst, 5, 5, 0.0, 0.0, s);
}

// YUYV -> BGR -> Gray -> f32 -> Sobel magnitude.  One fused strip kernel (strip_yuyv_sobel.cu) whenever the
// strip path applies; tiny or unaligned images run the chain's stand-alone kernels over device scratch.
int launch_yuyv_sobel_strip(Ctx *c, const DBatch &src, const DBatch &mag, cudaStream_t s);

int launch_yuyv_sobel(Ctx *c, const DBatch &src, const DBatch &mag, cudaStream_t s) {
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  if (opt_get("yuyvsobel.force_chain", 0) == 0) {
    int rc = launch_yuyv_sobel_strip(c, src, mag, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }
  if (mag.windowed()) return RCV_ERR_UNSUPPORTED;
  DBatch g8 = mag, g32 = mag;
  const size_t p8 = ((size_t)src.v.cols + 255) / 256 * 256, p32 = ((size_t)src.v.cols * 4 + 255) / 256 * 256;
  void *a = nullptr, *b = nullptr;
  RCV_TRY(ctx_scratch(c, SCR_FUSE_TMP, p8 * src.v.rows * src.n, &a));
  RCV_TRY(ctx_scratch(c, SCR_FUSE_TMP2, p32 * src.v.rows * src.n, &b));
  g8.v.data = (uint8_t *)a;
  g8.v.step = p8;
  g8.v.depth = RCV_U8;
  g8.frame_stride = src.n > 1 ? p8 * src.v.rows : 0;
  g32.v.data = (uint8_t *)b;
  g32.v.step = p32;
  g32.frame_stride = src.n > 1 ? p32 * src.v.rows : 0;
  RCV_TRY(launch_cvt(c, src, g8, RCV_COLOR_YUYV2GRAY, s));
  RCV_TRY(launch_convert(c, g8, g32, 1.0, 0.0, s));
  DBatch none;
  memset(&none, 0, sizeof(none));
  none.n = src.n;
  return launch_sobel(c, g32, mag, none, none, s);
}

// ---------------------------------------------------------------------------------------
// dense filter2D (correlation, anchor = centre, REFLECT_101)
//   acc = delta; acc = fmaf(k[i][j], p, acc) in row-major tap order
//   u8: saturate(rint(acc)) (round half even)
// ---------------------------------------------------------------------------------------
struct F2dArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *dst;
  size_t dstep, dfs;
  int rows, cols, cn;
  int kw, kh;
  const float *taps;  // device, kh*kw
  float delta;
};

// 9 outputs per lane, not 8: a lane's first column is (lane / cn) * 9 * cn + lane % cn, and with an odd 9 the lanes of
// a warp fall into distinct shared-memory banks for 1, 2 and 4 channels (2-way at worst for 3) where 8 gave 3- to 8-way
// conflicts on every read of the f32 tile (dense 5x5 f32 BGR 0.18 -> 0.24 of the roofline, u8 7x7 0.046 -> 0.059)
constexpr int kF2dTBY = 32, kF2dThreads = 256, kF2dN = 9;
constexpr int kF2dMaxRW = 32 * kF2dN + (kMaxTaps - 1) * 4;

// A lane computes the kF2dN outputs of one channel of a block of kF2dN pixels of one tile row: kernel row by kernel
// row (ascending), each a tap_chain over that source row -- per output exactly the oracle's row-major fmaf chain.
template <typename T, int KN>
__global__ void __launch_bounds__(kF2dThreads) k_filter2d(const F2dArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int kw = KN > 0 ? KN : a.kw, kh = KN > 0 ? KN : a.kh;
  const int rx = kw / 2, ry = kh / 2;
  const int lanes = 32 / a.cn * a.cn;
  const int TBX = lanes * kF2dN;
  const int RW = TBX + (kw - 1) * a.cn;
  const int RH = kF2dTBY + kh - 1;
  float *taps = (float *)smem;
  float *raw = (float *)(smem + (((size_t)kw * kh * 4 + 15) & ~(size_t)15));  // staged as f32: converted once, not per tap
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NWARP = kF2dThreads / 32;
  const int ncols = a.cols * a.cn;
  const int ex0 = blockIdx.x * TBX, y0 = blockIdx.y * kF2dTBY;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  uint8_t *dst = a.dst + (size_t)blockIdx.z * a.dfs;
  for (int i = threadIdx.x; i < kw * kh; i += kF2dThreads) taps[i] = a.taps[i];
  constexpr int NX = (kF2dMaxRW + 31) / 32;
  const bool interior_x = ex0 - rx * a.cn >= 0 && ex0 - rx * a.cn + RW <= ncols;
  int xoff[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    const int x = lane + 32 * i;
    xoff[i] = 0;
    if (x < RW) {
      const int ex = ex0 + x - rx * a.cn;
      if (interior_x) {
        xoff[i] = ex;
      } else {
        const int px = ex >= 0 ? ex / a.cn : -((-ex + a.cn - 1) / a.cn);
        xoff[i] = reflect101(px, a.cols) * a.cn + (ex - px * a.cn);
      }
    }
  }
  for (int r = warp; r < RH; r += NWARP) {
    const T *srow = (const T *)(src + (size_t)reflect101(y0 + r - ry, a.rows) * a.sstep);
    float *rrow = raw + r * RW;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int x = lane + 32 * i;
      if (x < RW) rrow[x] = (float)srow[xoff[i]];
    }
  }
  __syncthreads();
  if (lane >= lanes) return;
  const int cn = a.cn;
  const int x0 = lane / cn * (kF2dN * cn) + lane % cn;
  for (int r = warp; r < kF2dTBY; r += NWARP) {
    const int y = y0 + r;
    if (y >= a.rows) break;
    float acc[kF2dN];
#pragma unroll
    for (int m = 0; m < kF2dN; ++m) acc[m] = a.delta;
    for (int ki = 0; ki < kh; ++ki) {
      const float *p = raw + (r + ki) * RW + x0;
      tap_chain<kF2dN, KN>(acc, (const float *)(taps + ki * kw), kw, [&](int i) { return p[i * cn]; },
                           [](float k, float v, float c) { return fmaf(k, v, c); });
    }
#pragma unroll
    for (int m = 0; m < kF2dN; ++m) {
      const int ex = ex0 + x0 + m * cn;
      if (ex >= ncols) continue;
      if (sizeof(T) == 1) {
        int v = __float2int_rn(acc[m]);  // round half to even, saturating conversion
        ((uint8_t *)(dst + (size_t)y * a.dstep))[ex] = (uint8_t)min(max(v, 0), 255);
      } else {
        ((float *)(dst + (size_t)y * a.dstep))[ex] = acc[m];
      }
    }
  }
}

template <typename T, int KN>
static int launch_f2d_kn(const F2dArgs &a, int n, cudaStream_t s) {
  const int kw = KN > 0 ? KN : a.kw, kh = KN > 0 ? KN : a.kh;
  const int TBX = tile_lanes(a.cn) * kF2dN;
  const int RW = TBX + (kw - 1) * a.cn, RH = kF2dTBY + kh - 1;
  const size_t smem = (((size_t)kw * kh * 4 + 15) & ~(size_t)15) + (size_t)RW * RH * sizeof(float);
  auto kern = k_filter2d<T, KN>;
  static size_t attr_smem[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_smem[dev & 15] < smem) {
    RCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[dev & 15] = smem;
  }
  dim3 grid(ceil_div(a.cols * a.cn, TBX), ceil_div(a.rows, kF2dTBY), n);
  if (grid.y > 65535 || grid.z > 65535) return fail(RCV_ERR_UNSUPPORTED, "image too tall / batch too large");
  kern<<<grid, kF2dThreads, smem, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

template <typename T>
static int launch_f2d_t(const F2dArgs &a, int n, cudaStream_t s) {
  if (a.kw == a.kh && a.kw == 3) return launch_f2d_kn<T, 3>(a, n, s);
  if (a.kw == a.kh && a.kw == 5) return launch_f2d_kn<T, 5>(a, n, s);
  if (a.kw == a.kh && a.kw == 7) return launch_f2d_kn<T, 7>(a, n, s);
  return launch_f2d_kn<T, 0>(a, n, s);
}

int launch_filter2d(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                    cudaStream_t s) {
  if (kw < 1 || kh < 1 || kw > kMaxTaps || kh > kMaxTaps)
    return fail(RCV_ERR_ARG, "kernel size %dx%d outside 1..%d", kw, kh, kMaxTaps);
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  if (opt_get("f2d.force_generic", 0) == 0) {
    int rc = src.v.depth == RCV_F32 ? launch_filter2d_f32_strip(c, src, dst, k, kw, kh, delta, s)
                                    : launch_filter2d_u8_strip(c, src, dst, k, kw, kh, delta, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }
  if (dst.windowed()) return RCV_ERR_UNSUPPORTED;
  void *dtaps = nullptr;
  RCV_TRY(ctx_scratch(c, SCR_TAPS, (size_t)kw * kh * sizeof(float), &dtaps));
  RCV_CUDA(cudaMemcpyAsync(dtaps, k, (size_t)kw * kh * sizeof(float), cudaMemcpyHostToDevice, s));
  F2dArgs a{src.v.data, src.v.step, src.frame_stride, dst.v.data, dst.v.step, dst.frame_stride, src.v.rows,
            src.v.cols, src.v.cn, kw, kh, (const float *)dtaps, delta};
  return src.v.depth == RCV_U8 ? launch_f2d_t<uint8_t>(a, src.n, s) : launch_f2d_t<float>(a, src.n, s);
}

// ---------------------------------------------------------------------------------------
// Sobel 3x3 + magnitude, general geometry (one thread per pixel; index-arithmetic borders).
// Same operation order as Sobel3Op in stencil.cu and orc_sobel3_f32.
// ---------------------------------------------------------------------------------------
struct SobelArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *out[3];  // mag, gx, gy
  size_t ostep[3], ofs[3];
  int rows, cols;
};

__global__ void __launch_bounds__(256) k_sobel_generic(const SobelArgs a) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= a.cols) return;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const float *pm = (const float *)(src + (size_t)reflect101(y - 1, a.rows) * a.sstep);
  const float *p0 = (const float *)(src + (size_t)y * a.sstep);
  const float *pp = (const float *)(src + (size_t)reflect101(y + 1, a.rows) * a.sstep);
  int xs[3] = {reflect101(x - 1, a.cols), x, reflect101(x + 1, a.cols)};
  float s[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float t = __fadd_rn(pm[xs[k]], pp[xs[k]]);
    float u = __fmul_rn(2.0f, p0[xs[k]]);
    s[k] = __fadd_rn(t, u);
    d[k] = __fsub_rn(pp[xs[k]], pm[xs[k]]);
  }
  float gx = __fsub_rn(s[2], s[0]);
  float gy = __fadd_rn(__fadd_rn(d[0], d[2]), __fmul_rn(2.0f, d[1]));
  float mg = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
  float res[3] = {mg, gx, gy};
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (a.out[k]) ((float *)(a.out[k] + (size_t)blockIdx.z * a.ofs[k] + (size_t)y * a.ostep[k]))[x] = res[k];
}

int launch_sobel_strip(Ctx *c, const DBatch &src, const DBatch &mag, const DBatch &gx, const DBatch &gy,
                       cudaStream_t s);

int launch_sobel(Ctx *c, const DBatch &src, const DBatch &mag, const DBatch &gx, const DBatch &gy,
                 cudaStream_t s) {
  if (src.v.rows == 0 || src.v.cols == 0 || src.n == 0) return RCV_OK;
  if (opt_get("sobel.force_generic", 0) == 0) {
    int rc = launch_sobel_strip(c, src, mag, gx, gy, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }
  if (mag.windowed()) return RCV_ERR_UNSUPPORTED;
  if (src.v.rows > 65535 || src.n > 65535) return fail(RCV_ERR_UNSUPPORTED, "image too tall / batch too large");
  SobelArgs a;
  a.src = src.v.data;
  a.sstep = src.v.step;
  a.sfs = src.frame_stride;
  const DBatch *o[3] = {&mag, &gx, &gy};
  for (int k = 0; k < 3; ++k) {
    a.out[k] = o[k]->v.data;
    a.ostep[k] = o[k]->v.step;
    a.ofs[k] = o[k]->frame_stride;
  }
  a.rows = src.v.rows;
  a.cols = src.v.cols;
  dim3 grid(ceil_div(a.cols, 256), a.rows, src.n);
  k_sobel_generic<<<grid, 256, 0, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

}  // namespace rcv
