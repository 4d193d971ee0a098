/*
 * rcv_imgproc.h -- C ABI of librcv_imgproc.so, the B200 (sm_100a) imgproc
 * backend for RustCV.
 *
 * This is the drop-in boundary.  RustCV has no plugin registry for image ops:
 * they are free functions over `Mat` (rustcv/src/imgproc/drawing.rs:67,
 * rustcv/src/highgui/mod.rs:24), so the Rust side binds these entry points in
 * an `extern "C"` block exactly as rustcv-camera binds its one native bridge
 * (rustcv-camera/src/backend/macos/mod.rs:42-80 over
 * rustcv-camera/src/backend/macos/bridge.h:17-65).  The conventions below are
 * that bridge's conventions:
 *   - plain C types only, no C++/torch types, no exceptions across the ABI;
 *   - every function returns int: 0 = OK, negative = error (bridge.h:20-24);
 *   - results through out-pointers, caller-allocated buffers (bridge.h:36-62);
 *   - opaque handles with an explicit free (bridge.h:17,65);
 *   - the message for the last error on this thread: rcv_last_error().
 *
 * Threading contract: every call is synchronous with respect to the caller
 * (it returns after the work on the Mat is complete) unless
 * rcv_set_blocking(0) was called, re-entrant, and thread-safe for distinct
 * Mats.  One CUDA context per GPU; work for a Mat runs on the GPU named by
 * RcvMat.device (device Mats) or on the GPU given to rcv_init (host Mats).
 * The *_batch_multi calls fan a batch out over several GPUs from the ONE
 * calling thread (the reference API is a single synchronous caller,
 * README.md:31, rustcv/src/videoio/mod.rs:168): the library owns one worker
 * thread per GPU and returns when every frame is done.
 *
 * There is NO CPU fallback: without a B200-class GPU rcv_init returns
 * RCV_ERR_CUDA and every op returns RCV_ERR_NOT_INIT.
 */
#ifndef RCV_IMGPROC_H
#define RCV_IMGPROC_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define RCV_API __attribute__((visibility("default")))
#else
#define RCV_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes (bridge.h:20-24 convention) --------------------------- */
#define RCV_OK 0
#define RCV_ERR_ARG (-1)         /* NULL pointer, bad enum, bad kernel size   */
#define RCV_ERR_SIZE (-2)        /* dst geometry/step/length does not match   */
#define RCV_ERR_DEPTH (-3)       /* depth/channels not supported by this op   */
#define RCV_ERR_CUDA (-4)        /* CUDA runtime/driver failure               */
#define RCV_ERR_UNSUPPORTED (-5) /* valid request this build does not cover   */
#define RCV_ERR_NOT_INIT (-6)    /* rcv_init has not succeeded                */
#define RCV_ERR_NOMEM (-7)       /* device or pinned allocation failed        */
#define RCV_ERR_NCCL (-8)        /* the coefficient broadcast failed (NCCL)   */

/* ---- Mat --------------------------------------------------------------- */
/* depth tag: rustcv::core::Mat is u8-only (rustcv/src/core/mat.rs:6-15, TODO
 * at :53); f32 images keep `step` in BYTES and store floats in `data`. */
#define RCV_U8 0
#define RCV_F32 1

/* where `data` lives */
#define RCV_HOST 0        /* pageable host memory (a Rust Vec<u8>)            */
#define RCV_DEVICE 1      /* HBM, from rcv_mat_alloc_device                   */
#define RCV_HOST_PINNED 2 /* page-locked host memory from rcv_pinned_alloc    */

/* POD mirror of rustcv::core::Mat (rustcv/src/core/mat.rs:6-15): row r starts
 * at data + r*step, the first cols*channels*elemsize bytes of a row are valid
 * (row_bytes, mat.rs:47-51).  Caller-owned: the library never frees or
 * reallocates host memory; dst must arrive correctly sized. */
typedef struct RcvMat {
  void *data;
  int32_t rows;
  int32_t cols;
  size_t step;      /* bytes per row, >= cols*channels*elemsize */
  uint8_t channels; /* 1..4 */
  uint8_t depth;    /* RCV_U8 | RCV_F32 */
  uint8_t loc;      /* RCV_HOST | RCV_DEVICE | RCV_HOST_PINNED */
  uint8_t reserved;
  int32_t device;   /* GPU ordinal for RCV_DEVICE */
} RcvMat;

/* ---- lifecycle --------------------------------------------------------- */
/* Binds the calling process to GPU `device` (creates the context, streams and
 * staging rings).  May be called for several devices; the first one becomes
 * the default device for host Mats.  Fails with RCV_ERR_CUDA when the GPU is
 * not compute capability 10.x. */
RCV_API int rcv_init(int device);
/* rcv_init for GPUs 0..ngpus-1 (ngpus <= 0: every GPU of the box) and one
 * worker thread per GPU, bound to the CPUs local to it.  rcv_init(-1) binds
 * to the GPU named by the environment variable RCV_DEVICE (default 0). */
RCV_API int rcv_init_multi(int32_t ngpus);
RCV_API int rcv_shutdown(void);
RCV_API int rcv_device_count(int *count);
/* 1 (default): every op returns after its work completed.  0: ops on device
 * Mats only enqueue on the library stream; call rcv_sync() to wait. */
RCV_API int rcv_set_blocking(int blocking);
RCV_API int rcv_sync(int device);
/* The cudaStream_t the library launches on for `device` (for event timing). */
RCV_API int rcv_get_stream(int device, void **stream);
/* Number of kernels this library has launched since rcv_init. */
RCV_API int rcv_launch_count(uint64_t *count);
RCV_API const char *rcv_last_error(void);
RCV_API const char *rcv_version(void);

/* ---- storage ----------------------------------------------------------- */
/* Device-resident Mat storage: fills *m (data, step, loc, device) for the
 * given geometry.  step is rounded up to 256 B (the reference's own default
 * `align_stride: Some(256)`, rustcv-core/src/builder.rs:33) so rows are TMA-
 * and 128-bit-access aligned. */
RCV_API int rcv_mat_alloc_device(RcvMat *m, int32_t rows, int32_t cols, int32_t channels, int32_t depth, int32_t device);
RCV_API int rcv_mat_free_device(RcvMat *m);
/* n Mats of one geometry carved from ONE allocation (frame j at data + j*frame_bytes). */
RCV_API int rcv_mat_alloc_device_batch(RcvMat *mats, int32_t n, int32_t rows, int32_t cols, int32_t channels,
                               int32_t depth, int32_t device);
RCV_API int rcv_mat_free_device_batch(RcvMat *mats, int32_t n);
/* strided copies host<->device (re-pitching); geometry must match. */
RCV_API int rcv_mat_upload(const RcvMat *host, RcvMat *dev);
RCV_API int rcv_mat_download(const RcvMat *dev, RcvMat *host);
RCV_API int rcv_pinned_alloc(void **ptr, size_t bytes);
/* The same, with the pages placed on the NUMA node of GPU `device` (the GPU
 * that will DMA them) when the box exposes several nodes; device < 0 = the
 * default GPU.  Free with rcv_pinned_free. */
RCV_API int rcv_pinned_alloc_on(int32_t device, void **ptr, size_t bytes);
RCV_API int rcv_pinned_free(void *ptr);
/* Page-locks a CALLER-OWNED buffer in place -- the Vec<u8> behind a reference
 * Mat, which read() reuses frame after frame (rustcv/src/videoio/mod.rs:192-199,
 * rustcv-camera/src/mat.rs:65-74).  Afterwards an RCV_HOST Mat inside the range
 * is DMA'd directly, like RCV_HOST_PINNED.  The caller must unregister before
 * freeing or reallocating the buffer (the Rust wrapper does so in Drop and in
 * ensure_size).  Unregistered pageable Mats still work: they go
 * through the library's pinned bounce ring. */
RCV_API int rcv_host_register(void *ptr, size_t bytes);
RCV_API int rcv_host_unregister(void *ptr);

/* ---- pixel-format conversion (cvtColor) ---------------------------------
 * The reference's own hot loops: rustcv/src/videoio/mod.rs:344-399, twins in
 * rustcv-camera/src/decode.rs:160-228. */
#define RCV_COLOR_YUYV2BGR 0   /* src channels=2 (Y,U|V interleaved), dst 3 */
#define RCV_COLOR_UYVY2BGR 1   /* src channels=2, dst 3                     */
#define RCV_COLOR_BGRA2BGR 2   /* src 4, dst 3 (videoio/mod.rs:385-399)     */
#define RCV_COLOR_RGB2BGR 3    /* src 3, dst 3, swap 0<->2 (decode.rs:213)  */
#define RCV_COLOR_BGR2RGB 3
#define RCV_COLOR_BGR2GRAY 4   /* src 3, dst 1                              */
#define RCV_COLOR_BGR2XRGB32 5 /* src 3, dst 4: 0x00RRGGBB (highgui/mod.rs:125-141) */
#define RCV_COLOR_YUYV2GRAY 6  /* src 2, dst 1: BGR2GRAY(YUYV2BGR(.)) fused */

RCV_API int rcv_cvt_color(const RcvMat *src, RcvMat *dst, int32_t code);
/* YUYV -> BGR on Mats; replaces the call at rustcv/src/videoio/mod.rs:203 and
 * rustcv-camera/src/decode.rs:52.  Stride-aware: cols/2 macro-pixels per row. */
RCV_API int rcv_yuyv_to_bgr(const RcvMat *src, RcvMat *dst);
/* Flat-buffer form with the facade's exact contract (videoio/mod.rs:344-371):
 * packed input, width*height/2 macro-pixels, stride ignored.  Where the
 * reference silently returns (src short) or panics (dst short) this returns
 * RCV_ERR_SIZE and leaves dst untouched. */
RCV_API int rcv_yuyv_to_bgr_packed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len, size_t width,
                           size_t height);
/* videoio/mod.rs:385-399 on flat host buffers. */
RCV_API int rcv_bgra_to_bgr_packed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len, size_t width,
                           size_t height);
/* NV12 (rustcv-backend-msmf/examples/camera_view/convert.rs:46-86): y is
 * rows x cols x 1, uv is rows/2 x cols/2 x 2. */
RCV_API int rcv_nv12_to_bgr(const RcvMat *y, const RcvMat *uv, RcvMat *dst);

/* The MJPEG branch of read(): header, then decompress to BGR at dst's pitch -- the two TurboJPEG calls of
 * rustcv/src/videoio/mod.rs:205-232 (rustcv-camera/src/decode.rs:93-121), through nvJPEG.  dst is u8,
 * channels = 3, rows x cols as reported by rcv_mjpeg_info; a device dst keeps the frame in HBM (only the
 * compressed bytes cross PCIe).  Decoders differ in the last bits: parity with libjpeg-turbo is a tolerance. */
RCV_API int rcv_mjpeg_info(const uint8_t *jpeg, size_t len, int32_t *width, int32_t *height);
RCV_API int rcv_mjpeg_to_bgr(const uint8_t *jpeg, size_t len, RcvMat *dst);

/* cv::Mat::convertTo between u8 and f32 Mats of one geometry (the seam between the u8 images the
 * reference's capture path produces and the f32 images Sobel / warpAffine take):
 * v = fmaf((float)src, (float)alpha, (float)beta); u8 results are saturate(rint(v)). */
RCV_API int rcv_convert_to(const RcvMat *src, RcvMat *dst, double alpha, double beta);

/* ---- filtering (absent from the reference; OpenCV semantics, oracle/) --- */
/* cv::GaussianBlur, BORDER_REFLECT_101.  kw/kh odd (or 0 = derive from sigma);
 * sigma_y <= 0 means sigma_x.  u8: Q8 taps, single rounding; f32: fmaf chains. */
RCV_API int rcv_gaussian_blur(const RcvMat *src, RcvMat *dst, int32_t kw, int32_t kh, double sigma_x, double sigma_y);
/* separable filter, f32 images, f32 taps */
RCV_API int rcv_sep_filter2d(const RcvMat *src, RcvMat *dst, const float *kx, int32_t kw, const float *ky, int32_t kh);
/* separable filter, u8 images, Q8 integer taps: (sum ky kx p + 2^15) >> 16 */
RCV_API int rcv_sep_filter2d_q8(const RcvMat *src, RcvMat *dst, const int32_t *kx, int32_t kw, const int32_t *ky,
                        int32_t kh);
/* dense correlation, anchor at centre, row-major taps (kh rows of kw) */
RCV_API int rcv_filter2d(const RcvMat *src, RcvMat *dst, const float *kernel, int32_t kw, int32_t kh, float delta);
/* Sobel 3x3 on 1-channel f32; any of mag/gx/gy may be NULL (at least one set) */
RCV_API int rcv_sobel_mag(const RcvMat *src, RcvMat *mag, RcvMat *gx, RcvMat *gy);

/* ---- geometry ----------------------------------------------------------- */
/* cv::resize INTER_LINEAR; output size is dst's geometry. */
RCV_API int rcv_resize_bilinear(const RcvMat *src, RcvMat *dst);
/* cv::warpAffine INTER_LINEAR, BORDER_CONSTANT(border_value).  M is the
 * forward map unless inverse_map != 0. */
RCV_API int rcv_warp_affine(const RcvMat *src, RcvMat *dst, const double M[6], int32_t inverse_map,
                    double border_value);
RCV_API int rcv_get_rotation_matrix_2d(double cx, double cy, double angle_deg, double scale, double M[6]);
RCV_API int rcv_invert_affine(const double M[6], double iM[6]);

/* ---- fused chains (the step upstream of every imgproc call) ------------- */
/* GaussianBlur5x5(YUYV2BGR(src)) in ONE kernel (2 B/px in, 3 B/px out instead of
 * the two kernels' 11 B/px); bit-identical to rcv_yuyv_to_bgr + rcv_gaussian_blur. */
RCV_API int rcv_yuyv_to_bgr_gaussian5(const RcvMat *src_yuyv, RcvMat *dst_bgr);
/* SobelMagnitude(convertTo_f32(BGR2GRAY(YUYV2BGR(src)))) in ONE kernel: the raw
 * camera frame (rustcv/src/videoio/mod.rs:201-205, channels=2, even cols) in,
 * the f32 gradient magnitude (channels=1, same rows x cols) out -- 6 B/px of
 * HBM traffic instead of the chain's 22.  Bit-identical to the chain. */
RCV_API int rcv_yuyv_to_sobel_mag(const RcvMat *src_yuyv, RcvMat *mag_f32);

/* ---- batches of independent frames -------------------------------------
 * srcs[i] -> dsts[i], i < n, all of one geometry and location.  Device Mats:
 * ONE kernel launch for the whole batch.  Host Mats: H2D / kernel / D2H are
 * pipelined over the frames through the pinned staging ring. */
RCV_API int rcv_gaussian_blur_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t kw, int32_t kh,
                            double sigma_x, double sigma_y);
RCV_API int rcv_sobel_mag_batch(const RcvMat *srcs, RcvMat *mags, int32_t n);
RCV_API int rcv_resize_bilinear_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n);
RCV_API int rcv_warp_affine_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const double M[6],
                          int32_t inverse_map, double border_value);
RCV_API int rcv_cvt_color_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t code);
RCV_API int rcv_yuyv_to_sobel_mag_batch(const RcvMat *srcs_yuyv, RcvMat *mags_f32, int32_t n);
RCV_API int rcv_yuyv_to_bgr_gaussian5_batch(const RcvMat *srcs_yuyv, RcvMat *dsts_bgr, int32_t n);
RCV_API int rcv_sep_filter2d_q8_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const int32_t *kx, int32_t kw,
                              const int32_t *ky, int32_t kh);
/* dense filter2D (rcv_filter2d) over a batch: u8 or f32, kw x kh taps row-major, one launch for a uniform device batch */
RCV_API int rcv_filter2d_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const float *kernel, int32_t kw, int32_t kh,
                       float delta);

/* ---- the same batches sharded over several GPUs -----
 * Frames are independent: host Mats go frame j -> GPU j mod ngpus, device Mats
 * run on the GPU that owns them (src and dst of a frame on the same GPU).
 * ngpus <= 0: every initialised GPU (rcv_init_multi).  Each GPU runs its share
 * on its own worker thread, streams and staging ring; there is no inter-GPU
 * traffic on the pixel path.  Returns the first failing frame's code. */
RCV_API int rcv_gaussian_blur_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, int32_t kw,
                                  int32_t kh, double sigma_x, double sigma_y);
RCV_API int rcv_sobel_mag_batch_multi(const RcvMat *srcs, RcvMat *mags, int32_t n, int32_t ngpus);
RCV_API int rcv_resize_bilinear_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus);
RCV_API int rcv_warp_affine_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, const double M[6],
                                int32_t inverse_map, double border_value);
RCV_API int rcv_cvt_color_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, int32_t code);
RCV_API int rcv_yuyv_to_sobel_mag_batch_multi(const RcvMat *srcs_yuyv, RcvMat *mags_f32, int32_t n, int32_t ngpus);
RCV_API int rcv_yuyv_to_bgr_gaussian5_batch_multi(const RcvMat *srcs_yuyv, RcvMat *dsts_bgr, int32_t n, int32_t ngpus);
/* kx == NULL and ky == NULL: every GPU filters with the taps IT received from
 * rcv_set_kernel_broadcast (kw taps for x, then kh taps for y). */
RCV_API int rcv_sep_filter2d_q8_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus,
                                    const int32_t *kx, int32_t kw, const int32_t *ky, int32_t kh);
RCV_API int rcv_filter2d_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, const float *kernel,
                             int32_t kw, int32_t kh, float delta);

/* The path's single collective: filter coefficients are set up once, on GPU
 * `root_device`, and broadcast (ncclBroadcast over NVLink) into the coefficient
 * bank of the other initialised GPUs (ngpus <= 0: all).  `coeffs`: count <=
 * RCV_COEFF_BANK_MAX f32 values.  `received` (optional, ngpus x count floats)
 * returns each GPU's copy, read back from its bank.  RCV_ERR_NCCL on failure. */
#define RCV_COEFF_BANK_MAX 64
RCV_API int rcv_set_kernel_broadcast(const float *coeffs, int32_t count, int32_t root_device, int32_t ngpus,
                             float *received);

/* ---- tuning knobs (benchmark/diagnostic use) ----------------------------- */
/* name/value integer options, e.g. "gauss.band_rows", "gauss.variant". */
RCV_API int rcv_set_option(const char *name, int64_t value);
RCV_API int rcv_get_option(const char *name, int64_t *value);

#ifdef __cplusplus
}
#endif
#endif /* RCV_IMGPROC_H */
