// multi.cu -- several GPUs behind ONE synchronous caller.
//
// RustCV's API is one blocking call from one thread (README.md:31, rustcv/src/videoio/mod.rs:168); BASELINE.json's
// north_star shards "batches of independent frames one-frame-per-GPU ... with at most a single NCCL broadcast of
// filter coefficients and no inter-GPU traffic on the pixel path" (SURVEY.md section 8e).  So the fan-out lives
// here, inside the library: one worker thread per GPU (its own context, three streams and staging ring --
// context.cu), bound to the CPUs local to that GPU; a rcv_*_batch_multi call hands each worker its frames
// (frame j -> GPU j mod N for host Mats, the owning GPU for device Mats) and returns when all are done.
//
// The one collective, rcv_set_kernel_broadcast: the coefficients of a filter are set up once, on the root GPU, and
// ncclBroadcast to every other GPU's coefficient bank (one ncclComm per GPU from ncclCommInitAll, NVLink /
// NVSwitch underneath); the *_multi filter calls with NULL taps launch with what their GPU received.  NCCL is
// resolved at run time (libnccl.so.2, the copy a hosting process such as PyTorch already loaded if there is one):
// nothing else in the library needs it, and a process that never broadcasts never loads it.
#include "rcv_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

namespace rcv {

// ---------------------------------------------------------------------------------------------------------
// workers
// ---------------------------------------------------------------------------------------------------------
struct MultiJoin {
  std::mutex mu;
  std::condition_variable cv;
  int pending = 0;
  std::atomic<int> refs{0};  // the waiter + one per job: whoever lets go last frees it (never while a worker is still
                             // inside mu's unlock path)
  std::vector<int> rc;
  std::vector<std::string> err;
};

static void multi_release(MultiJoin *j) {
  if (j->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete j;
}

namespace {
struct Job {
  MultiJoin *join;
  int slot;
  int (*fn)(void *);
  void *arg;
};

struct Worker {
  int device = -1;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job> q;
  bool stop = false;

  void loop() {
    cudaSetDevice(device);
    bind_thread_to_gpu(device);
    char name[16];
    snprintf(name, sizeof(name), "rcv-gpu%d", device);
    pthread_setname_np(pthread_self(), name);
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return stop || !q.empty(); });
        if (q.empty()) return;
        j = q.front();
        q.pop_front();
      }
      const int rc = j.fn(j.arg);
      {
        std::lock_guard<std::mutex> lk(j.join->mu);
        j.join->rc[j.slot] = rc;
        if (rc != RCV_OK) j.join->err[j.slot] = last_error();
        if (--j.join->pending == 0) j.join->cv.notify_all();
      }
      multi_release(j.join);
    }
  }
};

constexpr int kMaxWorkers = 16;
Worker *g_workers[kMaxWorkers] = {};
std::mutex g_workers_mu;

Worker *worker_for(int device) {
  std::lock_guard<std::mutex> lk(g_workers_mu);
  if (device < 0 || device >= kMaxWorkers) return nullptr;
  if (!g_workers[device]) {
    Worker *w = new Worker();
    w->device = device;
    w->th = std::thread([w] { w->loop(); });
    g_workers[device] = w;
  }
  return g_workers[device];
}
}  // namespace

MultiJoin *multi_begin(int njobs) {
  MultiJoin *j = new MultiJoin();
  j->pending = njobs;
  j->refs.store(njobs + 1);
  j->rc.assign(njobs, RCV_OK);
  j->err.assign(njobs, std::string());
  return j;
}

void multi_submit(MultiJoin *j, int device, int slot, int (*fn)(void *), void *arg) {
  Worker *w = worker_for(device);
  if (!w) {
    {
      std::lock_guard<std::mutex> lk(j->mu);
      j->rc[slot] = RCV_ERR_ARG;
      j->err[slot] = "no worker for this device";
      if (--j->pending == 0) j->cv.notify_all();
    }
    multi_release(j);
    return;
  }
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->q.push_back(Job{j, slot, fn, arg});
  }
  w->cv.notify_one();
}

int multi_wait(MultiJoin *j) {
  int rc = RCV_OK;
  {
    std::unique_lock<std::mutex> lk(j->mu);
    j->cv.wait(lk, [j] { return j->pending == 0; });
    for (size_t i = 0; i < j->rc.size(); ++i)
      if (j->rc[i] != RCV_OK) {
        rc = j->rc[i];
        set_error("%s", j->err[i].c_str());
        break;
      }
  }
  multi_release(j);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::vector<int> devs;
  std::vector<ncclComm_t> comms;
};
Nccl g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  if (g_nccl.lib) return RCV_OK;
  const char *env = getenv("RCV_NCCL_LIB");
  void *h = nullptr;
  if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);  // the host process's copy, if loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(RCV_ERR_NCCL, "libnccl.so.2 not found (set RCV_NCCL_LIB): %s", dlerror());
  g_nccl.lib = h;
#define RCV_NCCL_SYM(field, name)                                         \
  *(void **)(&g_nccl.field) = dlsym(h, name);                            \
  if (!g_nccl.field) {                                                   \
    g_nccl.lib = nullptr;                                                \
    return fail(RCV_ERR_NCCL, "symbol %s missing from libnccl", name);   \
  }
  RCV_NCCL_SYM(CommInitAll, "ncclCommInitAll")
  RCV_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  RCV_NCCL_SYM(Broadcast, "ncclBroadcast")
  RCV_NCCL_SYM(GroupStart, "ncclGroupStart")
  RCV_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  RCV_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  RCV_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef RCV_NCCL_SYM
  return RCV_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
  return fail(RCV_ERR_NCCL, "NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
}
#define RCV_NCCL(expr)                                   \
  do {                                                   \
    ncclResult_t _r = (expr);                            \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr); \
  } while (0)

void nccl_teardown() {
  for (size_t i = 0; i < g_nccl.comms.size(); ++i)
    if (g_nccl.comms[i]) {
      cudaSetDevice(g_nccl.devs[i]);
      g_nccl.CommDestroy(g_nccl.comms[i]);
    }
  g_nccl.comms.clear();
  g_nccl.devs.clear();
}
}  // namespace

void multi_shutdown() {
  {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    nccl_teardown();
  }
  std::lock_guard<std::mutex> lk(g_workers_mu);
  for (int d = 0; d < kMaxWorkers; ++d) {
    Worker *w = g_workers[d];
    if (!w) continue;
    {
      std::lock_guard<std::mutex> lk2(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
    w->th.join();
    delete w;
    g_workers[d] = nullptr;
  }
}

}  // namespace rcv

using namespace rcv;

extern "C" {

int rcv_init_multi(int32_t ngpus) {
  int have = 0;
  RCV_CUDA(cudaGetDeviceCount(&have));
  if (ngpus <= 0) ngpus = have;
  if (ngpus > have) return fail(RCV_ERR_ARG, "rcv_init_multi(%d): the box has %d GPU(s)", ngpus, have);
  for (int d = 0; d < ngpus; ++d) RCV_TRY(rcv_init(d));
  for (int d = 0; d < ngpus; ++d)
    if (!worker_for(d)) return fail(RCV_ERR_ARG, "device ordinal %d beyond the worker table", d);
  return RCV_OK;
}

// The path's single collective.  `coeffs` = `count` f32 values (filter taps, or the six
// affine terms) in host memory, valid on the calling thread only; they are placed in the root GPU's coefficient
// bank and ncclBroadcast into the bank of every other initialised GPU (`ngpus` <= 0: all of them).  `received`
// (optional, count x ngpus floats) gets every GPU's copy read back from ITS bank, in device order -- what the
// *_multi filter calls launch with.
int rcv_set_kernel_broadcast(const float *coeffs, int32_t count, int32_t root_device, int32_t ngpus, float *received) {
  if (!coeffs || count < 1 || count > RCV_COEFF_BANK_MAX) return fail(RCV_ERR_ARG, "coeffs NULL or count %d outside 1..%d", count, RCV_COEFF_BANK_MAX);
  int devs[16];
  int n = devices_initialised(devs, 16);
  if (n == 0) return fail(RCV_ERR_NOT_INIT, "rcv_init has not been called (no CPU fallback exists)");
  if (ngpus > 0 && ngpus < n) n = ngpus;
  int root = -1;
  for (int i = 0; i < n; ++i)
    if (devs[i] == root_device) root = i;
  if (root < 0) return fail(RCV_ERR_ARG, "root device %d is not among the %d initialised GPU(s)", root_device, n);
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  Ctx *ctx[16];
  for (int i = 0; i < n; ++i) {
    ctx[i] = ctx_get(devs[i]);
    if (!ctx[i]) return RCV_ERR_NOT_INIT;
    if (!ctx[i]->coeff_bank) RCV_CUDA(cudaMalloc(&ctx[i]->coeff_bank, RCV_COEFF_BANK_MAX * sizeof(float)));
  }
  ctx_get(devs[root]);
  RCV_CUDA(cudaMemcpyAsync(ctx[root]->coeff_bank, coeffs, count * sizeof(float), cudaMemcpyHostToDevice, ctx[root]->stream));
  if (n > 1) {
    RCV_TRY(nccl_load());
    bool same = g_nccl.devs.size() == (size_t)n;
    for (int i = 0; same && i < n; ++i) same = g_nccl.devs[i] == devs[i];
    if (!same) {
      nccl_teardown();
      g_nccl.comms.assign(n, nullptr);
      g_nccl.devs.assign(devs, devs + n);
      ncclResult_t r = g_nccl.CommInitAll(g_nccl.comms.data(), n, devs);
      if (r != ncclSuccess) {
        g_nccl.comms.clear();
        g_nccl.devs.clear();
        return nccl_fail(r, "ncclCommInitAll");
      }
    }
    RCV_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
      ncclResult_t r = g_nccl.Broadcast(ctx[i]->coeff_bank, ctx[i]->coeff_bank, (size_t)count, ncclFloat32, root, g_nccl.comms[i],
                                        ctx[i]->stream);
      if (r != ncclSuccess) {
        g_nccl.GroupEnd();
        return nccl_fail(r, "ncclBroadcast");
      }
    }
    RCV_NCCL(g_nccl.GroupEnd());
  }
  for (int i = 0; i < n; ++i) {
    ctx_get(devs[i]);
    float tmp[RCV_COEFF_BANK_MAX];
    RCV_CUDA(cudaMemcpyAsync(tmp, ctx[i]->coeff_bank, count * sizeof(float), cudaMemcpyDeviceToHost, ctx[i]->stream));
    RCV_CUDA(cudaStreamSynchronize(ctx[i]->stream));
    {
      std::lock_guard<std::mutex> lk2(ctx[i]->mu);
      memcpy(ctx[i]->coeffs, tmp, count * sizeof(float));
      ctx[i]->n_coeffs = count;
    }
    if (received) memcpy(received + (size_t)i * count, tmp, count * sizeof(float));
  }
  return RCV_OK;
}

}  // extern "C"
