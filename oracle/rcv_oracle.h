/*
 * rcv_oracle.h -- CPU oracle for the rustcv imgproc hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and there only as the checker or as the
 * timed CPU arm.  The product (librcv_imgproc.so) never links or calls it.
 *
 * Parity status:
 *   - yuyv/bgra/rgb conversions: PINNED to the reference -- a line-by-line
 *     restatement of rustcv/src/videoio/mod.rs:344-399 and
 *     rustcv-camera/src/decode.rs:160-228, checked against the reference's own
 *     unit tests (decode.rs:234-273) in tests/test_oracle.py.
 *   - GaussianBlur / sepFilter / filter2D / Sobel / resize / warpAffine /
 *     BGR->Gray: "parity unpinned" by the reference (the ops are ABSENT from
 *     RustCV @07b07dd).  The reference advertises OpenCV
 *     parity (README.md:19,30), so the integer specs here were pinned against
 *     OpenCV 4.13 outputs instead (tests/golden/make_golden.py, bit-exact for
 *     the u8 ops); the f32 ops define their own operation order.
 *
 * All images are strided row buffers exactly like rustcv::core::Mat
 * (rustcv/src/core/mat.rs:6-15): row r starts at data + r*step, the first
 * cols*channels*elemsize bytes of a row are valid.
 */
#ifndef RCV_ORACLE_H
#define RCV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic data ------------------------------ */
uint64_t orc_splitmix64_next(uint64_t *state);
void orc_fill_u8(uint64_t seed, uint8_t *dst, size_t n);
void orc_fill_f32(uint64_t seed, float *dst, size_t n);
uint32_t orc_crc32(const uint8_t *p, size_t n);

/* number of worker threads used by every row-parallel function below
 * (1 = scalar port, the reference's own threading model). */
void orc_set_threads(int n);
int orc_get_threads(void);

/* ---- pixel-format conversion (the reference's real hot loops) ----------- */
/* returns 0 = converted, 1 = silently returned (short buffer, like the
 * reference), -1 = the reference would have panicked (dst too short). */
int orc_yuyv_to_bgr_facade(const uint8_t *src, size_t src_len, uint8_t *dst,
                           size_t dst_len, size_t width, size_t height);
int orc_yuyv_to_bgr_camera(const uint8_t *src, size_t src_len, uint8_t *dst,
                           size_t dst_len, size_t width, size_t height);
void orc_yuyv_to_bgr_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols);
void orc_uyvy_to_bgr_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols);
void orc_nv12_to_bgr_strided(const uint8_t *y, size_t ystep, const uint8_t *uv,
                             size_t uvstep, uint8_t *dst, size_t dstep,
                             int rows, int cols);
int orc_bgra_to_bgr_facade(const uint8_t *src, size_t src_len, uint8_t *dst,
                           size_t dst_len, size_t width, size_t height);
void orc_bgra_to_bgr_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols);
void orc_swap_rb_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                         size_t dstep, int rows, int cols);
void orc_bgr_to_gray_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                             size_t dstep, int rows, int cols);
void orc_bgr_to_xrgb32_strided(const uint8_t *src, size_t sstep, uint32_t *dst,
                               size_t dstep, int rows, int cols);
void orc_yuyv_to_gray_strided(const uint8_t *src, size_t sstep, uint8_t *dst,
                              size_t dstep, int rows, int cols);

/* cv::Mat::convertTo between u8 and f32 (depth codes 0 = u8, 1 = f32), any channel count:
 * v = fmaf((float)src, (float)alpha, (float)beta); u8 results are saturate(rint(v)) (half to even).
 * ncols = cols * channels. */
void orc_convert_to(const void *src, size_t sstep, int sdepth, void *dst, size_t dstep, int ddepth, int rows,
                    int ncols, double alpha, double beta);

/* ---- filtering ---------------------------------------------------------- */
/* OpenCV-compatible Gaussian taps.  kq sums to 256 (Q8), kd to ~1. */
int orc_gaussian_ksize(double sigma, int is_u8);
void orc_gaussian_kernel_f64(int n, double sigma, double *kd);
void orc_gaussian_kernel_q8(int n, double sigma, int *kq);

/* exact integer separable filter: (sum ky_i kx_j p + 2^15) >> 16, REFLECT_101 */
void orc_sepfilter_u8_q8(const uint8_t *src, size_t sstep, uint8_t *dst,
                         size_t dstep, int rows, int cols, int cn,
                         const int *kx, int kw, const int *ky, int kh);
void orc_gaussian_blur_u8(const uint8_t *src, size_t sstep, uint8_t *dst,
                          size_t dstep, int rows, int cols, int cn, int kw,
                          int kh, double sigma_x, double sigma_y);
/* tuned (auto-vectorising) restatement of the 5x5 sigma=0 case; bit-identical, used only as bench.py's second
 * CPU figure */
void orc_gauss5_binomial_u8_fast(const uint8_t *src, size_t sstep, uint8_t *dst, size_t dstep, int rows, int cols,
                                 int cn);
/* f32 separable: row pass then column pass, fmaf chains in ascending tap order */
void orc_sepfilter_f32(const float *src, size_t sstep, float *dst, size_t dstep,
                       int rows, int cols, int cn, const float *kx, int kw,
                       const float *ky, int kh);
void orc_gaussian_blur_f32(const float *src, size_t sstep, float *dst,
                           size_t dstep, int rows, int cols, int cn, int kw,
                           int kh, double sigma_x, double sigma_y);
/* dense correlation, anchor at centre, fmaf chain in row-major tap order */
void orc_filter2d_f32(const float *src, size_t sstep, float *dst, size_t dstep,
                      int rows, int cols, int cn, const float *k, int kw,
                      int kh, float delta);
void orc_filter2d_u8(const uint8_t *src, size_t sstep, uint8_t *dst,
                     size_t dstep, int rows, int cols, int cn, const float *k,
                     int kw, int kh, float delta);
/* Sobel 3x3 on single-channel f32; any of gx/gy/mag may be NULL */
void orc_sobel3_f32(const float *src, size_t sstep, float *gx, size_t gxstep,
                    float *gy, size_t gystep, float *mag, size_t magstep,
                    int rows, int cols);

/* ---- geometry ----------------------------------------------------------- */
void orc_resize_bilinear_u8(const uint8_t *src, size_t sstep, int srows,
                            int scols, uint8_t *dst, size_t dstep, int drows,
                            int dcols, int cn);
void orc_resize_bilinear_f32(const float *src, size_t sstep, int srows,
                             int scols, float *dst, size_t dstep, int drows,
                             int dcols, int cn);
void orc_rotation_matrix(double cx, double cy, double angle_deg, double scale,
                         double M[6]);
int orc_invert_affine(const double M[6], double iM[6]);
/* inverse-map bilinear, BORDER_CONSTANT(border_value); M is the forward map
 * unless inverse_map != 0.  Returns the number of source pixels touched via
 * *touched when non-NULL (for the algorithmic-bytes figure). */
void orc_warp_affine_f32(const float *src, size_t sstep, int srows, int scols,
                         float *dst, size_t dstep, int drows, int dcols,
                         const double M[6], int inverse_map, float border_value);
void orc_warp_affine_u8(const uint8_t *src, size_t sstep, int srows, int scols,
                        uint8_t *dst, size_t dstep, int drows, int dcols,
                        int cn, const double M[6], int inverse_map,
                        int border_value);

#ifdef __cplusplus
}
#endif
#endif
