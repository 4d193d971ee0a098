/*
 * rcv_oracle.c -- CPU oracle (plain C99 + pthreads).  TEST INFRASTRUCTURE ONLY,
 * see rcv_oracle.h for the rules and the parity status of each function.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, so
 * every f32 result is exactly the operation order written here).
 *
 * Reference citations are relative to /root/reference (RustCV @07b07dd).
 */
#include "rcv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* synthetic data: SplitMix64                          */
/* ------------------------------------------------------------------------ */
uint64_t orc_splitmix64_next(uint64_t *s) {
  uint64_t z;
  *s += 0x9E3779B97F4A7C15ULL;
  z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

void orc_fill_u8(uint64_t seed, uint8_t *dst, size_t n) {
  uint64_t s = seed;
  size_t i = 0;
  while (i < n) {
    uint64_t z = orc_splitmix64_next(&s);
    for (int b = 0; b < 8 && i < n; ++b, ++i) dst[i] = (uint8_t)(z >> (8 * b));
  }
}

void orc_fill_f32(uint64_t seed, float *dst, size_t n) {
  uint64_t s = seed;
  size_t i = 0;
  while (i < n) {
    uint64_t z = orc_splitmix64_next(&s);
    for (int h = 0; h < 2 && i < n; ++h, ++i) {
      uint32_t v = (uint32_t)(z >> (32 * h));
      dst[i] = (float)(v >> 8) * (1.0f / 16777216.0f);
    }
  }
}

uint32_t orc_crc32(const uint8_t *p, size_t n) {
  static uint32_t table[256];
  static int init = 0;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      table[i] = c;
    }
    init = 1;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

/* ------------------------------------------------------------------------ */
/* row-parallel helper                                                        */
/* ------------------------------------------------------------------------ */
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int orc_get_threads(void) { return g_threads; }

typedef void (*row_fn)(const void *args, int r0, int r1);
typedef struct {
  row_fn fn;
  const void *args;
  int r0, r1;
} job_t;

static void *job_main(void *p) {
  job_t *j = (job_t *)p;
  j->fn(j->args, j->r0, j->r1);
  return NULL;
}

static void parallel_rows(row_fn fn, const void *args, int rows) {
  int nt = g_threads;
  if (nt > rows) nt = rows;
  if (nt <= 1) {
    fn(args, 0, rows);
    return;
  }
  pthread_t tid[256];
  job_t jobs[256];
  for (int t = 0; t < nt; ++t) {
    jobs[t].fn = fn;
    jobs[t].args = args;
    jobs[t].r0 = (int)((long long)rows * t / nt);
    jobs[t].r1 = (int)((long long)rows * (t + 1) / nt);
    pthread_create(&tid[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nt; ++t) pthread_join(tid[t], NULL);
}

/* OpenCV borderInterpolate(BORDER_REFLECT_101): gfedcb|abcdefgh|gfedcba */
static int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0)
      p = -p;
    else
      p = 2 * (len - 1) - p;
  }
  return p;
}

/* ------------------------------------------------------------------------ */
/* pixel-format conversion                                                    */
/* ------------------------------------------------------------------------ */
/* rustcv/src/videoio/mod.rs:373-382 and rustcv-camera/src/decode.rs:225-228 */
static inline uint8_t clamp_u8(int32_t v) { return v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v); }

/* One macro-pixel [Y0,U,Y1,V] -> 2 BGR pixels.
 * rustcv/src/videoio/mod.rs:352-369 (identical math decode.rs:173-189).
 * i32 arithmetic; `>> 8` is an arithmetic (floor) shift on negative values,
 * as in Rust -- gcc implements signed >> as arithmetic. */
static inline void yuv_pair_to_bgr(int32_t y0, int32_t ub, int32_t y1, int32_t vb,
                                   uint8_t *d) {
  int32_t u = ub - 128, v = vb - 128;
  int32_t c0 = y0 - 16, c1 = y1 - 16;
  d[0] = clamp_u8((298 * c0 + 516 * u + 128) >> 8);
  d[1] = clamp_u8((298 * c0 - 100 * u - 208 * v + 128) >> 8);
  d[2] = clamp_u8((298 * c0 + 409 * v + 128) >> 8);
  d[3] = clamp_u8((298 * c1 + 516 * u + 128) >> 8);
  d[4] = clamp_u8((298 * c1 - 100 * u - 208 * v + 128) >> 8);
  d[5] = clamp_u8((298 * c1 + 409 * v + 128) >> 8);
}

/* rustcv/src/videoio/mod.rs:344-371: packed, stride ignored, w*h/2 iterations,
 * silent return when src is short, NO dst length check (Rust would panic). */
int orc_yuyv_to_bgr_facade(const uint8_t *src, size_t src_len, uint8_t *dst,
                           size_t dst_len, size_t width, size_t height) {
  size_t frame_len = width * height * 2;
  if (src_len < frame_len) return 1;
  size_t pairs = width * height / 2;
  if (dst_len < pairs * 6) return -1; /* reference: index-out-of-bounds panic */
  for (size_t i = 0; i < pairs; ++i)
    yuv_pair_to_bgr(src[i * 4], src[i * 4 + 1], src[i * 4 + 2], src[i * 4 + 3], dst + i * 6);
  return 0;
}

/* rustcv-camera/src/decode.rs:160-191: same loop, checks src AND dst. */
int orc_yuyv_to_bgr_camera(const uint8_t *src, size_t src_len, uint8_t *dst,
                           size_t dst_len, size_t width, size_t height) {
  size_t pairs = width * height / 2;
  if (src_len < pairs * 4 || dst_len < pairs * 6) return 1;
  for (size_t i = 0; i < pairs; ++i)
    yuv_pair_to_bgr(src[i * 4], src[i * 4 + 1], src[i * 4 + 2], src[i * 4 + 3], dst + i * 6);
  return 0;
}

typedef struct {
  const uint8_t *src;
  size_t sstep;
  uint8_t *dst;
  size_t dstep;
  int cols;
  const uint8_t *src2;
  size_t s2step;
} cvt_args;

/* Stride-aware row loop in the shape of
 * rustcv-backend-msmf/examples/camera_view/convert.rs:13-43 (cols/2 pairs per
 * row; an odd trailing pixel is left untouched). */
static void yuyv_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int i = 0; i < a->cols / 2; ++i)
      yuv_pair_to_bgr(s[i * 4], s[i * 4 + 1], s[i * 4 + 2], s[i * 4 + 3], d + i * 6);
  }
}
void orc_yuyv_to_bgr_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, dst, dstep, cols, NULL, 0};
  parallel_rows(yuyv_rows, &a, rows);
}

/* UYVY = [U,Y0,V,Y1] (FourCC in rustcv-core/src/pixel_format.rs:39-55); same
 * BT.601 integer formula as YUYV. */
static void uyvy_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int i = 0; i < a->cols / 2; ++i)
      yuv_pair_to_bgr(s[i * 4 + 1], s[i * 4], s[i * 4 + 3], s[i * 4 + 2], d + i * 6);
  }
}
void orc_uyvy_to_bgr_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, dst, dstep, cols, NULL, 0};
  parallel_rows(uyvy_rows, &a, rows);
}

/* NV12: rustcv-backend-msmf/examples/camera_view/convert.rs:46-86 (per-pixel
 * formula with uv_row=row/2, uv_col=col/2), written as BGR bytes. */
static void nv12_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *yrow = a->src + (size_t)r * a->sstep;
    const uint8_t *uvrow = a->src2 + (size_t)(r / 2) * a->s2step;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int c = 0; c < a->cols; ++c) {
      int32_t y = yrow[c];
      int32_t u = (int32_t)uvrow[(c / 2) * 2] - 128;
      int32_t v = (int32_t)uvrow[(c / 2) * 2 + 1] - 128;
      int32_t cc = y - 16;
      d[c * 3 + 0] = clamp_u8((298 * cc + 516 * u + 128) >> 8);
      d[c * 3 + 1] = clamp_u8((298 * cc - 100 * u - 208 * v + 128) >> 8);
      d[c * 3 + 2] = clamp_u8((298 * cc + 409 * v + 128) >> 8);
    }
  }
}
void orc_nv12_to_bgr_strided(const uint8_t *y, size_t ystep, const uint8_t *uv,
                             size_t uvstep, uint8_t *dst, size_t dstep,
                             int rows, int cols) {
  cvt_args a = {y, ystep, dst, dstep, cols, uv, uvstep};
  parallel_rows(nv12_rows, &a, rows);
}

/* rustcv/src/videoio/mod.rs:385-399: silent return when either is short. */
int orc_bgra_to_bgr_facade(const uint8_t *src, size_t src_len, uint8_t *dst,
                           size_t dst_len, size_t width, size_t height) {
  size_t n = width * height;
  if (src_len < n * 4 || dst_len < n * 3) return 1;
  for (size_t i = 0; i < n; ++i) {
    dst[i * 3 + 0] = src[i * 4 + 0];
    dst[i * 3 + 1] = src[i * 4 + 1];
    dst[i * 3 + 2] = src[i * 4 + 2];
  }
  return 0;
}

static void bgra_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int c = 0; c < a->cols; ++c) {
      d[c * 3 + 0] = s[c * 4 + 0];
      d[c * 3 + 1] = s[c * 4 + 1];
      d[c * 3 + 2] = s[c * 4 + 2];
    }
  }
}
void orc_bgra_to_bgr_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, dst, dstep, cols, NULL, 0};
  parallel_rows(bgra_rows, &a, rows);
}

/* rustcv-camera/src/decode.rs:213-219 (twins videoio/mod.rs:243-248,
 * imgcodecs/mod.rs:22-27,51-63): swap bytes 0 and 2. */
static void swaprb_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int c = 0; c < a->cols; ++c) {
      uint8_t b0 = s[c * 3 + 0], b1 = s[c * 3 + 1], b2 = s[c * 3 + 2];
      d[c * 3 + 0] = b2;
      d[c * 3 + 1] = b1;
      d[c * 3 + 2] = b0;
    }
  }
}
void orc_swap_rb_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                         size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, dst, dstep, cols, NULL, 0};
  parallel_rows(swaprb_rows, &a, rows);
}

/* OpenCV 4.13 cvtColor(BGR2GRAY) integer model:
 * (3735 B + 19235 G + 9798 R + 16384) >> 15.  Pinned vs cv2 in make_golden.py. */
static void gray_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int c = 0; c < a->cols; ++c)
      d[c] = (uint8_t)((3735u * s[c * 3] + 19235u * s[c * 3 + 1] + 9798u * s[c * 3 + 2] + 16384u) >> 15);
  }
}
void orc_bgr_to_gray_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, dst, dstep, cols, NULL, 0};
  parallel_rows(gray_rows, &a, rows);
}

/* rustcv/src/highgui/mod.rs:125-141: 0x00RRGGBB per pixel (stride-aware here;
 * the reference ignores step). */
static void xrgb_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint32_t *d = (uint32_t *)(a->dst + (size_t)r * a->dstep);
    for (int c = 0; c < a->cols; ++c)
      d[c] = ((uint32_t)s[c * 3 + 2] << 16) | ((uint32_t)s[c * 3 + 1] << 8) | s[c * 3];
  }
}
void orc_bgr_to_xrgb32_strided(const uint8_t *src, size_t sstep, uint32_t *dst,
                               size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, (uint8_t *)dst, dstep, cols, NULL, 0};
  parallel_rows(xrgb_rows, &a, rows);
}

/* YUYV -> Gray is defined as BGR2GRAY(yuyv_to_bgr(.)) so that the fused GPU
 * chain equals the two-step reference pipeline bit for bit. */
static void yuyv_gray_rows(const void *a_, int r0, int r1) {
  const cvt_args *a = (const cvt_args *)a_;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s = a->src + (size_t)r * a->sstep;
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int i = 0; i < a->cols / 2; ++i) {
      uint8_t bgr[6];
      yuv_pair_to_bgr(s[i * 4], s[i * 4 + 1], s[i * 4 + 2], s[i * 4 + 3], bgr);
      d[i * 2 + 0] = (uint8_t)((3735u * bgr[0] + 19235u * bgr[1] + 9798u * bgr[2] + 16384u) >> 15);
      d[i * 2 + 1] = (uint8_t)((3735u * bgr[3] + 19235u * bgr[4] + 9798u * bgr[5] + 16384u) >> 15);
    }
  }
}
void orc_yuyv_to_gray_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                              size_t dstep, int rows, int cols) {
  cvt_args a = {src, sstep, dst, dstep, cols, NULL, 0};
  parallel_rows(yuyv_gray_rows, &a, rows);
}

/* cv::Mat::convertTo model: one fmaf in f32, u8 results rounded half-to-even and saturated. */
void orc_convert_to(const void *src, size_t sstep, int sdepth, void *dst, size_t dstep, int ddepth, int rows,
                    int ncols, double alpha, double beta) {
  const float a = (float)alpha, b = (float)beta;
  for (int r = 0; r < rows; ++r) {
    const uint8_t *s8 = (const uint8_t *)src + (size_t)r * sstep;
    const float *sf = (const float *)s8;
    uint8_t *d8 = (uint8_t *)dst + (size_t)r * dstep;
    float *df = (float *)d8;
    for (int x = 0; x < ncols; ++x) {
      float v = fmaf(sdepth ? sf[x] : (float)s8[x], a, b);
      if (ddepth) {
        df[x] = v;
      } else {
        long iv = lrintf(v);
        d8[x] = iv < 0 ? 0 : (iv > 255 ? 255 : (uint8_t)iv);
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Gaussian taps (OpenCV model)                                               */
/* ------------------------------------------------------------------------ */
/* cv::GaussianBlur: ksize from sigma when ksize == 0. */
int orc_gaussian_ksize(double sigma, int is_u8) {
  int k = (int)lrint(sigma * (is_u8 ? 3 : 4) * 2 + 1);
  return k | 1;
}

/* cv::getGaussianKernel model: fixed table for sigma<=0 and n in {1,3,5,7},
 * otherwise exp(-x^2/(2 sigma^2)) normalised, sigma<=0 -> 0.3((n-1)/2-1)+0.8. */
void orc_gaussian_kernel_f64(int n, double sigma, double *kd) {
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

/* Q8 taps with OpenCV's error-diffusion rounding (outer taps inward, centre
 * takes the remainder so the taps sum to exactly 256). */
void orc_gaussian_kernel_q8(int n, double sigma, int *kq) {
  double kd[64];
  orc_gaussian_kernel_f64(n, sigma, kd);
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

/* ------------------------------------------------------------------------ */
/* exact u8 separable filter                                                  */
/* ------------------------------------------------------------------------ */
typedef struct {
  const uint8_t *src;
  size_t sstep;
  uint8_t *dst;
  size_t dstep;
  int rows, cols, cn;
  const int *kx;
  int kw;
  const int *ky;
  int kh;
} sepu8_args;

/* out = (sum_i sum_j ky[i] kx[j] p[r+i-ry][c+j-rx] + 32768) >> 16 with a SINGLE
 * rounding (intermediate rounding is wrong).  For the 5x5
 * sigma=0 taps {16,64,96,64,16} this is (sum k_i k_j p + 128) >> 8 with
 * k = {1,4,6,4,1}. */
static void sepu8_rows(const void *a_, int r0, int r1) {
  const sepu8_args *a = (const sepu8_args *)a_;
  int rx = a->kw / 2, ry = a->kh / 2;
  int n = a->cols * a->cn;
  uint32_t *hrow = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n * a->kh);
  int *xofs = (int *)malloc(sizeof(int) * (size_t)a->cols * a->kw);
  for (int c = 0; c < a->cols; ++c)
    for (int j = 0; j < a->kw; ++j) xofs[c * a->kw + j] = reflect101(c + j - rx, a->cols) * a->cn;
  /* ring of kh horizontally filtered rows, indexed by (source row) mod kh */
  int *have = (int *)malloc(sizeof(int) * a->kh);
  for (int i = 0; i < a->kh; ++i) have[i] = -1;
  for (int r = r0; r < r1; ++r) {
    for (int i = 0; i < a->kh; ++i) {
      int sr = reflect101(r + i - ry, a->rows);
      int slot = sr % a->kh;
      if (have[slot] == sr) continue;
      const uint8_t *s = a->src + (size_t)sr * a->sstep;
      uint32_t *h = hrow + (size_t)slot * n;
      for (int c = 0; c < a->cols; ++c)
        for (int ch = 0; ch < a->cn; ++ch) {
          uint32_t acc = 0;
          for (int j = 0; j < a->kw; ++j) acc += (uint32_t)a->kx[j] * s[xofs[c * a->kw + j] + ch];
          h[c * a->cn + ch] = acc;
        }
      have[slot] = sr;
    }
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int x = 0; x < n; ++x) {
      uint32_t acc = 32768u;
      for (int i = 0; i < a->kh; ++i) {
        int sr = reflect101(r + i - ry, a->rows);
        acc += (uint32_t)a->ky[i] * hrow[(size_t)(sr % a->kh) * n + x];
      }
      acc >>= 16;
      d[x] = acc > 255 ? 255 : (uint8_t)acc;
    }
  }
  free(have);
  free(xofs);
  free(hrow);
}

void orc_sepfilter_u8_q8(const uint8_t *src, size_t sstep, uint8_t *dst,
                         size_t dstep, int rows, int cols, int cn,
                         const int *kx, int kw, const int *ky, int kh) {
  sepu8_args a = {src, sstep, dst, dstep, rows, cols, cn, kx, kw, ky, kh};
  parallel_rows(sepu8_rows, &a, rows);
}

void orc_gaussian_blur_u8(const uint8_t *src, size_t sstep, uint8_t *dst,
                          size_t dstep, int rows, int cols, int cn, int kw,
                          int kh, double sigma_x, double sigma_y) {
  int kx[64], ky[64];
  if (sigma_y <= 0) sigma_y = sigma_x;
  if (kw <= 0 && sigma_x > 0) kw = orc_gaussian_ksize(sigma_x, 1);
  if (kh <= 0 && sigma_y > 0) kh = orc_gaussian_ksize(sigma_y, 1);
  orc_gaussian_kernel_q8(kw, sigma_x, kx);
  orc_gaussian_kernel_q8(kh, sigma_y, ky);
  orc_sepfilter_u8_q8(src, sstep, dst, dstep, rows, cols, cn, kx, kw, ky, kh);
}

/* The same 5x5 binomial GaussianBlur (ksize 5, sigma 0), written the way a tuned CPU library writes it: vertical
 * pass first into u16 (V <= 4080), then a branch-free horizontal pass over a row padded with its REFLECT_101
 * columns (H <= 65280 fits u16), single rounding -- every loop auto-vectorises.  Bit-identical to
 * orc_gaussian_blur_u8(.., 5, 5, 0, 0) (tests/test_oracle.py).  Not the definition: bench.py reports it beside the
 * definition port so that the CPU figure is not an artefact of scalar code. */
typedef struct {
  const uint8_t *src;
  size_t sstep;
  uint8_t *dst;
  size_t dstep;
  int rows, cols, cn;
} g5fast_args;

static void g5fast_rows(const void *a_, int r0, int r1) {
  const g5fast_args *a = (const g5fast_args *)a_;
  const int cn = a->cn, n = a->cols * cn, pad = 2 * cn;
  uint16_t *vp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(n + 2 * pad));
  uint16_t *v = vp + pad;
  for (int r = r0; r < r1; ++r) {
    const uint8_t *s0 = a->src + (size_t)reflect101(r - 2, a->rows) * a->sstep;
    const uint8_t *s1 = a->src + (size_t)reflect101(r - 1, a->rows) * a->sstep;
    const uint8_t *s2 = a->src + (size_t)r * a->sstep;
    const uint8_t *s3 = a->src + (size_t)reflect101(r + 1, a->rows) * a->sstep;
    const uint8_t *s4 = a->src + (size_t)reflect101(r + 2, a->rows) * a->sstep;
    for (int x = 0; x < n; ++x)
      v[x] = (uint16_t)(s0[x] + s4[x] + 4 * (s1[x] + s3[x]) + 6 * s2[x]);
    for (int k = 1; k <= 2; ++k) /* REFLECT_101 columns -k and cols-1+k */
      for (int ch = 0; ch < cn; ++ch) {
        v[-k * cn + ch] = v[reflect101(-k, a->cols) * cn + ch];
        v[(a->cols - 1 + k) * cn + ch] = v[reflect101(a->cols - 1 + k, a->cols) * cn + ch];
      }
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int x = 0; x < n; ++x) {
      uint16_t h = (uint16_t)(v[x - 2 * cn] + v[x + 2 * cn] + 4 * (v[x - cn] + v[x + cn]) + 6 * v[x] + 128);
      d[x] = (uint8_t)(h >> 8);
    }
  }
  free(vp);
}

void orc_gauss5_binomial_u8_fast(const uint8_t *src, size_t sstep, uint8_t *dst, size_t dstep, int rows, int cols,
                                 int cn) {
  g5fast_args a = {src, sstep, dst, dstep, rows, cols, cn};
  parallel_rows(g5fast_rows, &a, rows);
}

/* ------------------------------------------------------------------------ */
/* f32 separable filter                                                       */
/* ------------------------------------------------------------------------ */
typedef struct {
  const float *src;
  size_t sstep;
  float *dst;
  size_t dstep;
  int rows, cols, cn;
  const float *kx;
  int kw;
  const float *ky;
  int kh;
  float delta;
} sepf_args;

#define ROWF(base, step, r) ((const float *)((const uint8_t *)(base) + (size_t)(r) * (step)))
#define ROWFM(base, step, r) ((float *)((uint8_t *)(base) + (size_t)(r) * (step)))

/* row pass: h = fmaf(kx[j], p_j, h) for j ascending from h = 0;
 * column pass: v = fmaf(ky[i], h_i, v) for i ascending from v = 0. */
static void sepf_rows(const void *a_, int r0, int r1) {
  const sepf_args *a = (const sepf_args *)a_;
  int rx = a->kw / 2, ry = a->kh / 2;
  int n = a->cols * a->cn;
  float *hrow = (float *)malloc(sizeof(float) * (size_t)n * a->kh);
  int *xofs = (int *)malloc(sizeof(int) * (size_t)a->cols * a->kw);
  int *have = (int *)malloc(sizeof(int) * a->kh);
  for (int c = 0; c < a->cols; ++c)
    for (int j = 0; j < a->kw; ++j) xofs[c * a->kw + j] = reflect101(c + j - rx, a->cols) * a->cn;
  for (int i = 0; i < a->kh; ++i) have[i] = -1;
  for (int r = r0; r < r1; ++r) {
    for (int i = 0; i < a->kh; ++i) {
      int sr = reflect101(r + i - ry, a->rows);
      int slot = sr % a->kh;
      if (have[slot] == sr) continue;
      const float *s = ROWF(a->src, a->sstep, sr);
      float *h = hrow + (size_t)slot * n;
      for (int c = 0; c < a->cols; ++c)
        for (int ch = 0; ch < a->cn; ++ch) {
          float acc = 0.0f;
          for (int j = 0; j < a->kw; ++j) acc = fmaf(a->kx[j], s[xofs[c * a->kw + j] + ch], acc);
          h[c * a->cn + ch] = acc;
        }
      have[slot] = sr;
    }
    float *d = ROWFM(a->dst, a->dstep, r);
    for (int x = 0; x < n; ++x) {
      float acc = 0.0f;
      for (int i = 0; i < a->kh; ++i) {
        int sr = reflect101(r + i - ry, a->rows);
        acc = fmaf(a->ky[i], hrow[(size_t)(sr % a->kh) * n + x], acc);
      }
      d[x] = acc;
    }
  }
  free(have);
  free(xofs);
  free(hrow);
}

void orc_sepfilter_f32(const float *src, size_t sstep, float *dst, size_t dstep,
                       int rows, int cols, int cn, const float *kx, int kw,
                       const float *ky, int kh) {
  sepf_args a = {src, sstep, dst, dstep, rows, cols, cn, kx, kw, ky, kh, 0.0f};
  parallel_rows(sepf_rows, &a, rows);
}

void orc_gaussian_blur_f32(const float *src, size_t sstep, float *dst,
                           size_t dstep, int rows, int cols, int cn, int kw,
                           int kh, double sigma_x, double sigma_y) {
  double kd[64];
  float kx[64], ky[64];
  if (sigma_y <= 0) sigma_y = sigma_x;
  if (kw <= 0 && sigma_x > 0) kw = orc_gaussian_ksize(sigma_x, 0);
  if (kh <= 0 && sigma_y > 0) kh = orc_gaussian_ksize(sigma_y, 0);
  orc_gaussian_kernel_f64(kw, sigma_x, kd);
  for (int i = 0; i < kw; ++i) kx[i] = (float)kd[i];
  orc_gaussian_kernel_f64(kh, sigma_y, kd);
  for (int i = 0; i < kh; ++i) ky[i] = (float)kd[i];
  orc_sepfilter_f32(src, sstep, dst, dstep, rows, cols, cn, kx, kw, ky, kh);
}

/* ------------------------------------------------------------------------ */
/* dense filter2D (correlation, anchor = centre, REFLECT_101)                 */
/* ------------------------------------------------------------------------ */
/* acc = delta; acc = fmaf(k[i][j], p, acc) in row-major tap order. */
static void f2d_f32_rows(const void *a_, int r0, int r1) {
  const sepf_args *a = (const sepf_args *)a_;
  int rx = a->kw / 2, ry = a->kh / 2;
  for (int r = r0; r < r1; ++r) {
    float *d = ROWFM(a->dst, a->dstep, r);
    for (int c = 0; c < a->cols; ++c)
      for (int ch = 0; ch < a->cn; ++ch) {
        float acc = a->delta;
        for (int i = 0; i < a->kh; ++i) {
          const float *s = ROWF(a->src, a->sstep, reflect101(r + i - ry, a->rows));
          for (int j = 0; j < a->kw; ++j)
            acc = fmaf(a->kx[i * a->kw + j], s[reflect101(c + j - rx, a->cols) * a->cn + ch], acc);
        }
        d[c * a->cn + ch] = acc;
      }
  }
}
void orc_filter2d_f32(const float *src, size_t sstep, float *dst, size_t dstep,
                      int rows, int cols, int cn, const float *k, int kw,
                      int kh, float delta) {
  sepf_args a = {src, sstep, dst, dstep, rows, cols, cn, k, kw, NULL, kh, delta};
  parallel_rows(f2d_f32_rows, &a, rows);
}

typedef struct {
  const uint8_t *src;
  size_t sstep;
  uint8_t *dst;
  size_t dstep;
  int rows, cols, cn;
  const float *k;
  int kw, kh;
  float delta;
} f2du8_args;

/* same chain on (float)p; result = saturate_u8(lrintf(acc)) (round half even) */
static void f2d_u8_rows(const void *a_, int r0, int r1) {
  const f2du8_args *a = (const f2du8_args *)a_;
  int rx = a->kw / 2, ry = a->kh / 2;
  for (int r = r0; r < r1; ++r) {
    uint8_t *d = a->dst + (size_t)r * a->dstep;
    for (int c = 0; c < a->cols; ++c)
      for (int ch = 0; ch < a->cn; ++ch) {
        float acc = a->delta;
        for (int i = 0; i < a->kh; ++i) {
          const uint8_t *s = a->src + (size_t)reflect101(r + i - ry, a->rows) * a->sstep;
          for (int j = 0; j < a->kw; ++j)
            acc = fmaf(a->k[i * a->kw + j], (float)s[reflect101(c + j - rx, a->cols) * a->cn + ch], acc);
        }
        long v = lrintf(acc);
        d[c * a->cn + ch] = v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v);
      }
  }
}
void orc_filter2d_u8(const uint8_t *src, size_t sstep, uint8_t *dst,
                     size_t dstep, int rows, int cols, int cn, const float *k,
                     int kw, int kh, float delta) {
  f2du8_args a = {src, sstep, dst, dstep, rows, cols, cn, k, kw, kh, delta};
  parallel_rows(f2d_u8_rows, &a, rows);
}

/* ------------------------------------------------------------------------ */
/* Sobel 3x3 + magnitude (f32, single channel)                                */
/* ------------------------------------------------------------------------ */
typedef struct {
  const float *src;
  size_t sstep;
  float *gx;
  size_t gxstep;
  float *gy;
  size_t gystep;
  float *mag;
  size_t magstep;
  int rows, cols;
} sobel_args;

/* Separable form, every op a single correctly rounded f32 op (no fma):
 *   column smooth  s[c] = (p[r-1][c] + p[r+1][c]) + 2*p[r][c]
 *   column diff    d[c] =  p[r+1][c] - p[r-1][c]
 *   gx = s[c+1] - s[c-1]
 *   gy = (d[c-1] + d[c+1]) + 2*d[c]
 *   mag = sqrtf(gx*gx + gy*gy)
 * REFLECT_101 on rows and columns. */
static void sobel_rows(const void *a_, int r0, int r1) {
  const sobel_args *a = (const sobel_args *)a_;
  int n = a->cols;
  float *s = (float *)malloc(sizeof(float) * (size_t)n);
  float *d = (float *)malloc(sizeof(float) * (size_t)n);
  for (int r = r0; r < r1; ++r) {
    const float *pm = ROWF(a->src, a->sstep, reflect101(r - 1, a->rows));
    const float *p0 = ROWF(a->src, a->sstep, r);
    const float *pp = ROWF(a->src, a->sstep, reflect101(r + 1, a->rows));
    for (int c = 0; c < n; ++c) {
      float t = pm[c] + pp[c];
      float u = 2.0f * p0[c];
      s[c] = t + u;
      d[c] = pp[c] - pm[c];
    }
    float *ogx = a->gx ? ROWFM(a->gx, a->gxstep, r) : NULL;
    float *ogy = a->gy ? ROWFM(a->gy, a->gystep, r) : NULL;
    float *om = a->mag ? ROWFM(a->mag, a->magstep, r) : NULL;
    for (int c = 0; c < n; ++c) {
      int cl = reflect101(c - 1, n), cr = reflect101(c + 1, n);
      float gx = s[cr] - s[cl];
      float t = d[cl] + d[cr];
      float u = 2.0f * d[c];
      float gy = t + u;
      if (ogx) ogx[c] = gx;
      if (ogy) ogy[c] = gy;
      if (om) {
        float xx = gx * gx;
        float yy = gy * gy;
        om[c] = sqrtf(xx + yy);
      }
    }
  }
  free(s);
  free(d);
}
void orc_sobel3_f32(const float *src, size_t sstep, float *gx, size_t gxstep,
                    float *gy, size_t gystep, float *mag, size_t magstep,
                    int rows, int cols) {
  sobel_args a = {src, sstep, gx, gxstep, gy, gystep, mag, magstep, rows, cols};
  parallel_rows(sobel_rows, &a, rows);
}

/* ------------------------------------------------------------------------ */
/* bilinear resize (OpenCV INTER_LINEAR model, half-pixel centres)            */
/* ------------------------------------------------------------------------ */
typedef struct {
  const void *src;
  size_t sstep;
  int srows, scols;
  void *dst;
  size_t dstep;
  int drows, dcols, cn;
} resize_args;

/* cv::resize coordinate model: f = (float)((d+0.5)*scale - 0.5); s=floor(f);
 * f -= s.  Horizontal taps clamp to [0, scols-1] with f forced to 0 when
 * clamped; vertical rows clamp without touching f.  Fixed point for u8:
 * weights = rint(w * 2048) as int16 (round-half-even), horizontal sum in int,
 * vertical ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2. */
static inline void resize_coord_x(int dx, double scale, int slen, int *s0, float *f) {
  float fx = (float)((dx + 0.5) * scale - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) {
    fx = 0.0f;
    sx = 0;
  }
  if (sx >= slen - 1) {
    fx = 0.0f;
    sx = slen - 1;
  }
  *s0 = sx;
  *f = fx;
}

static inline int clipi(int x, int lo, int hi_excl) { return x < lo ? lo : (x >= hi_excl ? hi_excl - 1 : x); }

static void resize_u8_rows(const void *a_, int r0, int r1) {
  const resize_args *a = (const resize_args *)a_;
  double scale_x = (double)a->scols / a->dcols;
  double scale_y = (double)a->srows / a->drows;
  int *sx0 = (int *)malloc(sizeof(int) * a->dcols);
  short *ax = (short *)malloc(sizeof(short) * 2 * a->dcols);
  for (int dx = 0; dx < a->dcols; ++dx) {
    float fx;
    resize_coord_x(dx, scale_x, a->scols, &sx0[dx], &fx);
    ax[dx * 2 + 0] = (short)lrintf((1.0f - fx) * 2048.0f);
    ax[dx * 2 + 1] = (short)lrintf(fx * 2048.0f);
  }
  for (int dy = r0; dy < r1; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= (float)sy;
    int b0 = (short)lrintf((1.0f - fy) * 2048.0f);
    int b1 = (short)lrintf(fy * 2048.0f);
    const uint8_t *s0 = (const uint8_t *)a->src + (size_t)clipi(sy, 0, a->srows) * a->sstep;
    const uint8_t *s1 = (const uint8_t *)a->src + (size_t)clipi(sy + 1, 0, a->srows) * a->sstep;
    uint8_t *d = (uint8_t *)a->dst + (size_t)dy * a->dstep;
    for (int dx = 0; dx < a->dcols; ++dx) {
      int x0 = sx0[dx];
      int x1 = x0 + 1 < a->scols ? x0 + 1 : x0;
      int a0 = ax[dx * 2], a1 = ax[dx * 2 + 1];
      for (int ch = 0; ch < a->cn; ++ch) {
        int S0 = s0[x0 * a->cn + ch] * a0 + s0[x1 * a->cn + ch] * a1;
        int S1 = s1[x0 * a->cn + ch] * a0 + s1[x1 * a->cn + ch] * a1;
        int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        d[dx * a->cn + ch] = v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v);
      }
    }
  }
  free(sx0);
  free(ax);
}
void orc_resize_bilinear_u8(const uint8_t *src, size_t sstep, int srows,
                            int scols, uint8_t *dst, size_t dstep, int drows,
                            int dcols, int cn) {
  resize_args a = {src, sstep, srows, scols, dst, dstep, drows, dcols, cn};
  parallel_rows(resize_u8_rows, &a, drows);
}

/* f32: h = p0*(1-fx) + p1*fx ; out = h0*(1-fy) + h1*fy, each op rounded once. */
static void resize_f32_rows(const void *a_, int r0, int r1) {
  const resize_args *a = (const resize_args *)a_;
  double scale_x = (double)a->scols / a->dcols;
  double scale_y = (double)a->srows / a->drows;
  for (int dy = r0; dy < r1; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= (float)sy;
    float b0 = 1.0f - fy, b1 = fy;
    const float *s0 = ROWF(a->src, a->sstep, clipi(sy, 0, a->srows));
    const float *s1 = ROWF(a->src, a->sstep, clipi(sy + 1, 0, a->srows));
    float *d = ROWFM(a->dst, a->dstep, dy);
    for (int dx = 0; dx < a->dcols; ++dx) {
      int x0;
      float fx;
      resize_coord_x(dx, scale_x, a->scols, &x0, &fx);
      int x1 = x0 + 1 < a->scols ? x0 + 1 : x0;
      float a0 = 1.0f - fx, a1 = fx;
      for (int ch = 0; ch < a->cn; ++ch) {
        float t0 = s0[x0 * a->cn + ch] * a0;
        float t1 = s0[x1 * a->cn + ch] * a1;
        float h0 = t0 + t1;
        float t2 = s1[x0 * a->cn + ch] * a0;
        float t3 = s1[x1 * a->cn + ch] * a1;
        float h1 = t2 + t3;
        float v0 = h0 * b0;
        float v1 = h1 * b1;
        d[dx * a->cn + ch] = v0 + v1;
      }
    }
  }
}
void orc_resize_bilinear_f32(const float *src, size_t sstep, int srows,
                             int scols, float *dst, size_t dstep, int drows,
                             int dcols, int cn) {
  resize_args a = {src, sstep, srows, scols, dst, dstep, drows, dcols, cn};
  parallel_rows(resize_f32_rows, &a, drows);
}

/* ------------------------------------------------------------------------ */
/* warpAffine                                                                 */
/* ------------------------------------------------------------------------ */
/* cv::getRotationMatrix2D in f64. */
void orc_rotation_matrix(double cx, double cy, double angle_deg, double scale,
                         double M[6]) {
  double ang = angle_deg * (3.14159265358979323846 / 180.0);
  double alpha = scale * cos(ang);
  double beta = scale * sin(ang);
  M[0] = alpha;
  M[1] = beta;
  M[2] = (1 - alpha) * cx - beta * cy;
  M[3] = -beta;
  M[4] = alpha;
  M[5] = beta * cx + (1 - alpha) * cy;
}

/* cv::invertAffineTransform in f64. */
int orc_invert_affine(const double M[6], double iM[6]) {
  double D = M[0] * M[4] - M[1] * M[3];
  if (D == 0.0) return -1;
  D = 1.0 / D;
  double A11 = M[4] * D, A22 = M[0] * D;
  double A12 = -M[1] * D, A21 = -M[3] * D;
  double b1 = -A11 * M[2] - A12 * M[5];
  double b2 = -A21 * M[2] - A22 * M[5];
  iM[0] = A11;
  iM[1] = A12;
  iM[2] = b1;
  iM[3] = A21;
  iM[4] = A22;
  iM[5] = b2;
  return 0;
}

typedef struct {
  const void *src;
  size_t sstep;
  int srows, scols;
  void *dst;
  size_t dstep;
  int drows, dcols, cn;
  double iM[6];
  float border;
} warp_args;

/* Source coordinate of dst pixel (x, y), exactly:
 *   bx = (float)(iM[1]*y + iM[2])   (f64 mul, f64 add, one rounding to f32)
 *   sx = fmaf((float)iM[0], (float)x, bx)         likewise sy with iM[3..5]
 * then ix = floorf(sx), fx = sx - ix (exact), four taps with out-of-image taps
 * replaced by the border value, and the lerp chain
 *   r0 = fmaf(fx, p01 - p00, p00); r1 = fmaf(fx, p11 - p10, p10);
 *   out = fmaf(fy, r1 - r0, r0). */
static inline void warp_coord(const double *iM, int x, int y, float *sx, float *sy) {
  float bx = (float)(iM[1] * (double)y + iM[2]);
  float by = (float)(iM[4] * (double)y + iM[5]);
  *sx = fmaf((float)iM[0], (float)x, bx);
  *sy = fmaf((float)iM[3], (float)x, by);
}

static void warp_f32_rows(const void *a_, int r0, int r1) {
  const warp_args *a = (const warp_args *)a_;
  for (int y = r0; y < r1; ++y) {
    float *d = ROWFM(a->dst, a->dstep, y);
    for (int x = 0; x < a->dcols; ++x) {
      float sx, sy;
      warp_coord(a->iM, x, y, &sx, &sy);
      float flx = floorf(sx), fly = floorf(sy);
      float fx = sx - flx, fy = sy - fly;
      float p00 = a->border, p01 = a->border, p10 = a->border, p11 = a->border;
      /* reject coordinates that cannot touch the image before the int cast */
      if (flx >= -1.0f && flx < (float)a->scols && fly >= -1.0f && fly < (float)a->srows) {
        int ix = (int)flx, iy = (int)fly;
        int x0ok = ix >= 0, x1ok = ix + 1 < a->scols;
        if (iy >= 0) {
          const float *s = ROWF(a->src, a->sstep, iy);
          if (x0ok) p00 = s[ix];
          if (x1ok) p01 = s[ix + 1];
        }
        if (iy + 1 < a->srows) {
          const float *s = ROWF(a->src, a->sstep, iy + 1);
          if (x0ok) p10 = s[ix];
          if (x1ok) p11 = s[ix + 1];
        }
      }
      float q0 = fmaf(fx, p01 - p00, p00);
      float q1 = fmaf(fx, p11 - p10, p10);
      d[x] = fmaf(fy, q1 - q0, q0);
    }
  }
}

void orc_warp_affine_f32(const float *src, size_t sstep, int srows, int scols,
                         float *dst, size_t dstep, int drows, int dcols,
                         const double M[6], int inverse_map, float border_value) {
  warp_args a = {src, sstep, srows, scols, dst, dstep, drows, dcols, 1, {0}, border_value};
  if (inverse_map)
    memcpy(a.iM, M, sizeof(double) * 6);
  else if (orc_invert_affine(M, a.iM) != 0)
    return;
  parallel_rows(warp_f32_rows, &a, drows);
}

/* u8: same coordinates and lerp chain on (float)p per channel, result
 * saturate_u8(lrintf(v)). */
static void warp_u8_rows(const void *a_, int r0, int r1) {
  const warp_args *a = (const warp_args *)a_;
  int cn = a->cn;
  for (int y = r0; y < r1; ++y) {
    uint8_t *d = (uint8_t *)a->dst + (size_t)y * a->dstep;
    for (int x = 0; x < a->dcols; ++x) {
      float sx, sy;
      warp_coord(a->iM, x, y, &sx, &sy);
      float flx = floorf(sx), fly = floorf(sy);
      float fx = sx - flx, fy = sy - fly;
      int inside = flx >= -1.0f && flx < (float)a->scols && fly >= -1.0f && fly < (float)a->srows;
      int ix = inside ? (int)flx : 0, iy = inside ? (int)fly : 0;
      for (int ch = 0; ch < cn; ++ch) {
        float p00 = a->border, p01 = a->border, p10 = a->border, p11 = a->border;
        if (inside) {
          int x0ok = ix >= 0, x1ok = ix + 1 < a->scols;
          if (iy >= 0) {
            const uint8_t *s = (const uint8_t *)a->src + (size_t)iy * a->sstep;
            if (x0ok) p00 = (float)s[ix * cn + ch];
            if (x1ok) p01 = (float)s[(ix + 1) * cn + ch];
          }
          if (iy + 1 < a->srows) {
            const uint8_t *s = (const uint8_t *)a->src + (size_t)(iy + 1) * a->sstep;
            if (x0ok) p10 = (float)s[ix * cn + ch];
            if (x1ok) p11 = (float)s[(ix + 1) * cn + ch];
          }
        }
        float q0 = fmaf(fx, p01 - p00, p00);
        float q1 = fmaf(fx, p11 - p10, p10);
        long v = lrintf(fmaf(fy, q1 - q0, q0));
        d[x * cn + ch] = v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v);
      }
    }
  }
}

void orc_warp_affine_u8(const uint8_t *src, size_t sstep, int srows, int scols,
                        uint8_t *dst, size_t dstep, int drows, int dcols,
                        int cn, const double M[6], int inverse_map,
                        int border_value) {
  warp_args a = {src, sstep, srows, scols, dst, dstep, drows, dcols, cn, {0}, (float)border_value};
  if (inverse_map)
    memcpy(a.iM, M, sizeof(double) * 6);
  else if (orc_invert_affine(M, a.iM) != 0)
    return;
  parallel_rows(warp_u8_rows, &a, drows);
}
