// rcv_internal.cuh -- shared declarations of librcv_imgproc.so (not installed).
//
// Layout of the library:
//   context.cu   per-GPU context, streams, staging rings, device Mats, options
//   tma.cu       cuTensorMapEncodeTiled plumbing (driver entry point, no -lcuda)
//   cvt.cu       pixel-format conversion kernels     (videoio/mod.rs:344-399)
//   strip_*.cu   TMA strip-pipeline kernels (strip_pipeline.cuh): Gaussian u8, Sobel f32, filters, fused YUYV->Sobel
//   filter.cu    generic separable / dense filters (u8 Q8, f32)
//   geom.cu      bilinear resize, warpAffine
//   mjpeg.cu     the MJPEG branch of read() through nvJPEG (library decoder)
//   hostmem.cu   host-side memory: NUMA-placed pinned allocations, the registration cache for caller-owned
//                Vec<u8> buffers, the parallel memcpy pool behind the pageable bounce pipeline
//   multi.cu     one worker thread per GPU (frame j -> GPU j mod N inside ONE caller process), the NCCL
//                broadcast of filter coefficients
//   abi.cu       the extern "C" entry points of include/rcv_imgproc.h
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <atomic>
#include <mutex>
#include <set>

#include "../../include/rcv_imgproc.h"

namespace rcv {

// ---- errors ---------------------------------------------------------------
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define RCV_CUDA(expr)                                     \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return rcv::cuda_fail(_e, #expr); \
  } while (0)

#define RCV_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != RCV_OK) return _rc; \
  } while (0)

// ---- device image view -----------------------------------------------------
// A Mat whose storage is in HBM.  step in bytes; cn channels; depth RCV_U8/F32.
struct DView {
  uint8_t *data;
  int rows, cols;
  size_t step;
  int cn;
  int depth;
  __host__ __device__ size_t elem() const { return depth == RCV_F32 ? 4 : 1; }
  __host__ __device__ size_t row_bytes() const { return (size_t)cols * cn * elem(); }
};

// n frames of one geometry; frame j = data + j*frame_stride.
struct DBatch {
  DView v;              // geometry + frame 0
  size_t frame_stride;  // bytes between frames (0 when n == 1)
  int n;
  // Row window (on the DESTINATION batch of a launch): produce only output rows [y0, y1) of the image; source
  // rows outside the window's halo need not be resident yet.  y1 == 0 means all rows.  Used by the host
  // pipeline (abi.cu) to overlap H2D / kernel / D2H band by band inside ONE frame.  The strip launchers honour
  // it; a dispatcher that would have to fall back to a whole-image kernel returns RCV_ERR_UNSUPPORTED instead.
  int y0 = 0, y1 = 0;
  bool windowed() const { return y1 > 0; }
};

// ---- context ----------------------------------------------------------------
constexpr int kRing = 4;       // staging ring depth for host-resident batches
constexpr int kMaxBands = 32;  // row bands per frame in the pageable bounce pipeline (one event each)
constexpr int kTmapCache = 8;  // tensor maps remembered per context (one per recent (base, geometry))

enum ScratchSlot {
  SCR_STAGE_IN0 = 0,   // + slot            (kRing staged inputs; NV12 uses slot 1 for the UV plane)
  SCR_STAGE_OUT0 = 8,  // + slot*3 + output (kRing x up to 3 staged outputs)
  SCR_JPEG_Y = 20,     // + plane: Y, Cb, Cr samples of the MJPEG branch (20..22)
  SCR_TABLE_X = 24,
  SCR_TABLE_Y = 25,
  SCR_TAPS = 26,
  SCR_COUNTER = 28,   // work-stealing counter of the strip kernels
  SCR_FUSE_TMP = 27,  // intermediate BGR of the two-pass YUYV->BGR->Gaussian chain (gray u8 of the Sobel chain's backstop)
  SCR_FUSE_TMP2 = 29, // gray f32 of the unfused YUYV->Sobel backstop
  SCR_COUNT = 32
};

struct Ctx {
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;    // kernels (and single-Mat copies)
  cudaStream_t s_in = nullptr;      // H2D of pipelined host batches
  cudaStream_t s_out = nullptr;     // D2H of pipelined host batches
  cudaEvent_t ev_in[kRing] = {}, ev_k[kRing] = {}, ev_out[kRing] = {};
  void *scratch[SCR_COUNT] = {};
  size_t scratch_bytes[SCR_COUNT] = {};
  unsigned counter_parity = 0;       // which of the two strip-kernel work counters the next launch uses
  int resize_key[4] = {0, 0, 0, 0};  // geometry of the resize tables currently in SCR_TABLE_X/Y
  void *jpeg = nullptr;              // nvJPEG handle + state of the MJPEG branch (mjpeg.cu), created on first use
  // pinned bounce buffers of the pageable pipeline (abi.cu): one whole-frame buffer per ring slot and side,
  // grown on demand; ev_band[k][b] marks "band b of slot k has landed in bounce_out[k]".
  void *bounce_in[kRing] = {}, *bounce_out[kRing] = {};
  size_t bounce_in_bytes[kRing] = {}, bounce_out_bytes[kRing] = {};
  cudaEvent_t ev_band[kRing][kMaxBands] = {};
  // CUtensorMap cache: encoding a map costs a driver call per launch; the reference API is one call per frame
  // on the SAME reused buffers (rustcv/src/videoio/mod.rs:192-199), so the map of the last few
  // (base, geometry) pairs is kept.
  struct TmapEntry {
    const void *base = nullptr;
    size_t row_bytes = 0, step = 0, frame_stride = 0;
    int rows = 0, n = 0, box_w = 0, box_h = 0;
    uint64_t stamp = 0;
    CUtensorMap map;
  } tmaps[kTmapCache];
  uint64_t tmap_clock = 0;
  int numa_node = -1;                // NUMA node of the GPU (sysfs), -1 unknown / single node
  float *coeff_bank = nullptr;       // device copy of the broadcast filter coefficients (multi.cu)
  float coeffs[RCV_COEFF_BANK_MAX] = {};  // what this GPU received (read back from ITS bank after the broadcast)
  int n_coeffs = 0;
  std::atomic<int> out_pending[kRing];    // bounce_out[k] bands not yet copied to their Mat by the drain thread
  std::atomic<int> drain_failed{0};
  std::set<void *> dev_allocs;            // bases returned by rcv_mat_alloc_device[_batch] (what cudaFree may take)
  std::mutex mu;
};

Ctx *ctx_get(int device);      // NULL (and error set) when not initialised
Ctx *ctx_default();
inline cudaStream_t ctx_stream(Ctx *c) { return c->stream; }
inline int ctx_device(Ctx *c) { return c->device; }
inline int ctx_sm_count(Ctx *c) { return c->sm_count; }
bool ctx_blocking();
void count_launch(int n = 1);
int64_t opt_get(const char *name, int64_t dflt);

// device scratch that lives as long as the context (grown on demand).
int ctx_scratch(Ctx *c, int slot, size_t bytes, void **ptr);

// ---- TMA ---------------------------------------------------------------------
// 3-D tensor map over a batch of strided byte rows viewed as u32 words:
// dims {ceil(row_bytes/4), rows, n}, strides {step, frame_stride}, box {box_w, box_h, 1}.
int make_tmap_rows_u32(CUtensorMap *out, const void *base, size_t row_bytes, int rows, size_t step, int n,
                       size_t frame_stride, int box_w_words, int box_h);
// the same through the context's cache (c->mu held by the caller)
int ctx_tmap_rows_u32(Ctx *c, const CUtensorMap **out, const void *base, size_t row_bytes, int rows, size_t step, int n,
                      size_t frame_stride, int box_w_words, int box_h);

// ---- host memory (hostmem.cu) ----------------------------------------------------------------
// true when [p, p+bytes) lies inside memory this library pinned (rcv_pinned_alloc[_on]) or registered
// (rcv_host_register / the "host.auto_register" cache): DMA-able without a bounce copy.
bool host_range_pinned(const void *p, size_t bytes);
// "host.auto_register": register [p, p+bytes) on first sight (LRU over a bounded set); false = leave pageable
bool host_auto_register(const void *p, size_t bytes);
void host_registry_shutdown();
// rows x row_bytes strided copy on the pool's threads (the caller takes part); returns when done
void host_copy2d(void *dst, size_t dstep, const void *src, size_t sstep, size_t row_bytes, int rows);
// a band of a pageable destination: when `ev` (recorded after the band's D2H into the pinned bounce buffer) has
// fired, the GPU's drain thread copies rows x row_bytes to the Mat and decrements *pending
struct DrainJob {
  cudaEvent_t ev;
  void *dst;
  size_t dstep;
  const void *src;
  size_t sstep, row_bytes;
  int rows;
  std::atomic<int> *pending, *failed;
};
void drain_submit(int device, const DrainJob &job);
void drain_shutdown();
int gpu_numa_node(int device);                 // -1 when unknown
void bind_thread_to_gpu(int device);           // sched_setaffinity to the GPU's local_cpulist (if narrower)
const char *last_error();

// ---- multi-GPU workers (multi.cu) --------------------------------------------------------------
// Runs fn on the worker thread of `device` (started on first use); rc and error text come back through *rc / err.
struct MultiJoin;
MultiJoin *multi_begin(int njobs);
void multi_submit(MultiJoin *j, int device, int slot, int (*fn)(void *), void *arg);
int multi_wait(MultiJoin *j);  // first failing rc (its message becomes this thread's last error); frees j
void multi_shutdown();
int devices_initialised(int *out, int cap);  // ordinals with a live context, ascending

// ---- kernel launchers (device views only; enqueue on `s`) ------------------------
int launch_cvt(Ctx *c, const DBatch &src, const DBatch &dst, int code, cudaStream_t s);
int launch_nv12(Ctx *c, const DView &y, const DView &uv, const DView &dst, cudaStream_t s);
int launch_convert(Ctx *c, const DBatch &src, const DBatch &dst, double alpha, double beta, cudaStream_t s);

int launch_gaussian(Ctx *c, const DBatch &src, const DBatch &dst, int kw, int kh, double sx, double sy,
                    cudaStream_t s);
int launch_sepfilter_q8(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, int kw,
                        const int32_t *ky, int kh, cudaStream_t s);
int launch_sepfilter_f32(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky,
                         int kh, cudaStream_t s);
int launch_filter2d(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                    cudaStream_t s);
// any of mag/gx/gy may have data == NULL
int launch_sobel(Ctx *c, const DBatch &src, const DBatch &mag, const DBatch &gx, const DBatch &gy,
                 cudaStream_t s);
int launch_resize(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s);
int launch_warp_affine(Ctx *c, const DBatch &src, const DBatch &dst, const double iM[6], double border,
                       cudaStream_t s);
int launch_yuyv_gauss5(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s);
int launch_yuyv_sobel(Ctx *c, const DBatch &src, const DBatch &mag, cudaStream_t s);
// MJPEG branch (nvJPEG): header size, decode to BGR at dst.step, teardown
int mjpeg_info(Ctx *c, const uint8_t *jpeg, size_t len, int *width, int *height);
int launch_mjpeg(Ctx *c, const uint8_t *jpeg, size_t len, const DView &dst, cudaStream_t s);
void jpeg_destroy(Ctx *c);

// host-side helpers shared by abi.cu and the launchers (OpenCV models, see oracle/)
int gaussian_ksize(double sigma, bool is_u8);
void gaussian_kernel_f64(int n, double sigma, double *kd);
void gaussian_kernel_q8(int n, double sigma, int32_t *kq);
void rotation_matrix(double cx, double cy, double angle_deg, double scale, double M[6]);
int invert_affine(const double M[6], double iM[6]);

// ---- small device helpers ------------------------------------------------------------
__host__ __device__ inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace rcv
