// context.cu -- per-GPU context: streams, scratch, pinned staging ring, options.
//
// Replaces nothing in the reference one-to-one: RustCV's Mats are plain
// Vec<u8> (rustcv/src/core/mat.rs:6-15).  The device-resident storage variant and
// the pinned staging ring are the additions BASELINE.json's north_star asks for;
// the lifecycle follows the reference's only native bridge (open/free pairs,
// rustcv-camera/src/backend/macos/bridge.h:36-65).
#include "rcv_internal.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace rcv {

// ---- errors -------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char *what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  cudaGetLastError();  // clear the sticky-less error state
  return e == cudaErrorMemoryAllocation ? RCV_ERR_NOMEM : RCV_ERR_CUDA;
}

const char *last_error() { return g_err; }

// ---- context --------------------------------------------------------------------
static const int kMaxDevices = 16;

static Ctx *g_ctx[kMaxDevices] = {};
static int g_default_device = -1;
static std::mutex g_mu;
static std::atomic<int> g_blocking{1};
static std::atomic<uint64_t> g_launches{0};
static std::map<std::string, int64_t> g_opts;
static std::mutex g_opt_mu;
static std::atomic<int> g_opts_set{0};  // no option was ever set (every production call): opt_get takes no lock

Ctx *ctx_get(int device) {
  if (device < 0 || device >= kMaxDevices || !g_ctx[device]) {
    set_error("rcv_init(%d) has not been called (no CPU fallback exists)", device);
    return nullptr;
  }
  cudaSetDevice(device);
  return g_ctx[device];
}

Ctx *ctx_default() {
  if (g_default_device < 0) {
    set_error("rcv_init has not been called (no CPU fallback exists)");
    return nullptr;
  }
  return ctx_get(g_default_device);
}

bool ctx_blocking() { return g_blocking.load() != 0; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

int64_t opt_get(const char *name, int64_t dflt) {
  if (g_opts_set.load(std::memory_order_acquire) == 0) return dflt;
  std::lock_guard<std::mutex> lk(g_opt_mu);
  auto it = g_opts.find(name);
  return it == g_opts.end() ? dflt : it->second;
}

int ctx_scratch(Ctx *c, int slot, size_t bytes, void **ptr) {
  if (slot < 0 || slot >= SCR_COUNT) return fail(RCV_ERR_ARG, "bad scratch slot %d", slot);
  if (c->scratch_bytes[slot] < bytes) {
    // the stream may still be reading the old buffer
    RCV_CUDA(cudaStreamSynchronize(c->stream));
    if (c->scratch[slot]) RCV_CUDA(cudaFree(c->scratch[slot]));
    c->scratch[slot] = nullptr;
    c->scratch_bytes[slot] = 0;
    size_t want = bytes + bytes / 8 + 256;
    RCV_CUDA(cudaMalloc(&c->scratch[slot], want));
    c->scratch_bytes[slot] = want;
    if (slot == SCR_TABLE_X || slot == SCR_TABLE_Y) memset(c->resize_key, 0, sizeof(c->resize_key));
  }
  *ptr = c->scratch[slot];
  return RCV_OK;
}

int devices_initialised(int *out, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  int n = 0;
  for (int d = 0; d < kMaxDevices && n < cap; ++d)
    if (g_ctx[d]) out[n++] = d;
  return n;
}

int ctx_tmap_rows_u32(Ctx *c, const CUtensorMap **out, const void *base, size_t row_bytes, int rows, size_t step, int n,
                      size_t frame_stride, int box_w_words, int box_h) {
  Ctx::TmapEntry *victim = &c->tmaps[0];
  for (int i = 0; i < kTmapCache; ++i) {
    Ctx::TmapEntry &e = c->tmaps[i];
    if (e.base == base && e.row_bytes == row_bytes && e.rows == rows && e.step == step && e.n == n &&
        e.frame_stride == frame_stride && e.box_w == box_w_words && e.box_h == box_h) {
      e.stamp = ++c->tmap_clock;
      *out = &e.map;
      return RCV_OK;
    }
    if (e.stamp < victim->stamp) victim = &e;
  }
  RCV_TRY(make_tmap_rows_u32(&victim->map, base, row_bytes, rows, step, n, frame_stride, box_w_words, box_h));
  victim->base = base;
  victim->row_bytes = row_bytes;
  victim->rows = rows;
  victim->step = step;
  victim->n = n;
  victim->frame_stride = frame_stride;
  victim->box_w = box_w_words;
  victim->box_h = box_h;
  victim->stamp = ++c->tmap_clock;
  *out = &victim->map;
  return RCV_OK;
}

static int ctx_create(int device) {
  int n = 0;
  RCV_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n || device >= kMaxDevices)
    return fail(RCV_ERR_ARG, "device %d out of range (have %d)", device, n);
  cudaDeviceProp prop;
  RCV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(RCV_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                prop.minor);
  RCV_CUDA(cudaSetDevice(device));
  Ctx *c = new Ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
  for (int i = 0; i < kRing && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < kRing && e == cudaSuccess; ++i) {
    c->out_pending[i].store(0);
    for (int b = 0; b < kMaxBands && e == cudaSuccess; ++b)
      e = cudaEventCreateWithFlags(&c->ev_band[i][b], cudaEventDisableTiming | cudaEventBlockingSync);
  }
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(e, "stream/event creation");
  }
  c->numa_node = gpu_numa_node(device);
  g_ctx[device] = c;
  if (g_default_device < 0) g_default_device = device;
  return RCV_OK;
}

}  // namespace rcv

using namespace rcv;

extern "C" {

int rcv_init(int device) {
  // device < 0: the ordinal named by the environment variable RCV_DEVICE, else GPU 0 -- the library's one
  // piece of run-time configuration; like the reference there is no runtime backend dispatch
  // (rustcv/src/videoio/backend.rs:13-37 selects its backend at compile time).
  if (device < 0) {
    const char *env = getenv("RCV_DEVICE");
    device = (env && *env) ? atoi(env) : 0;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (device >= 0 && device < kMaxDevices && g_ctx[device]) return RCV_OK;
  return ctx_create(device);
}

int rcv_shutdown(void) {
  multi_shutdown();  // workers first: none may hold a context while it is torn down
  drain_shutdown();
  std::lock_guard<std::mutex> lk(g_mu);
  for (int d = 0; d < kMaxDevices; ++d) {
    Ctx *c = g_ctx[d];
    if (!c) continue;
    cudaSetDevice(d);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->s_in);
    cudaStreamSynchronize(c->s_out);
    jpeg_destroy(c);
    for (int i = 0; i < SCR_COUNT; ++i)
      if (c->scratch[i]) cudaFree(c->scratch[i]);
    for (int i = 0; i < kRing; ++i) {
      if (c->bounce_in[i]) cudaFreeHost(c->bounce_in[i]);
      if (c->bounce_out[i]) cudaFreeHost(c->bounce_out[i]);
      for (int b = 0; b < kMaxBands; ++b) cudaEventDestroy(c->ev_band[i][b]);
    }
    if (c->coeff_bank) cudaFree(c->coeff_bank);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->s_in);
    cudaStreamDestroy(c->s_out);
    for (int i = 0; i < kRing; ++i) {
      cudaEventDestroy(c->ev_in[i]);
      cudaEventDestroy(c->ev_k[i]);
      cudaEventDestroy(c->ev_out[i]);
    }
    delete c;
    g_ctx[d] = nullptr;
  }
  g_default_device = -1;
  host_registry_shutdown();
  cudaGetLastError();
  return RCV_OK;
}

int rcv_device_count(int *count) {
  if (!count) return fail(RCV_ERR_ARG, "count is NULL");
  RCV_CUDA(cudaGetDeviceCount(count));
  return RCV_OK;
}

int rcv_set_blocking(int blocking) {
  g_blocking.store(blocking ? 1 : 0);
  return RCV_OK;
}

int rcv_sync(int device) {
  Ctx *c = device < 0 ? ctx_default() : ctx_get(device);
  if (!c) return RCV_ERR_NOT_INIT;
  RCV_CUDA(cudaStreamSynchronize(c->stream));
  return RCV_OK;
}

int rcv_get_stream(int device, void **stream) {
  if (!stream) return fail(RCV_ERR_ARG, "stream is NULL");
  Ctx *c = device < 0 ? ctx_default() : ctx_get(device);
  if (!c) return RCV_ERR_NOT_INIT;
  *stream = (void *)c->stream;
  return RCV_OK;
}

int rcv_launch_count(uint64_t *count) {
  if (!count) return fail(RCV_ERR_ARG, "count is NULL");
  *count = g_launches.load();
  return RCV_OK;
}

const char *rcv_last_error(void) { return rcv::last_error(); }

const char *rcv_version(void) { return "rcv_imgproc 0.2 (sm_100a)"; }

int rcv_set_option(const char *name, int64_t value) {
  if (!name) return fail(RCV_ERR_ARG, "name is NULL");
  std::lock_guard<std::mutex> lk(g_opt_mu);
  g_opts[name] = value;
  g_opts_set.store(1, std::memory_order_release);
  return RCV_OK;
}

int rcv_get_option(const char *name, int64_t *value) {
  if (!name || !value) return fail(RCV_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(g_opt_mu);
  auto it = g_opts.find(name);
  if (it == g_opts.end()) return fail(RCV_ERR_ARG, "option %s is not set", name);
  *value = it->second;
  return RCV_OK;
}

}  // extern "C"
