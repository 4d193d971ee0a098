// hostmem.cu -- the host side of a Mat: where its bytes live and how they reach the copy engines.
//
// The reference Mat is a pageable Vec<u8> that `read()` REUSES frame after frame (rustcv/src/core/mat.rs:6-15,
// rustcv/src/videoio/mod.rs:192-199; rustcv-camera/src/mat.rs:65-74 `ensure_size`).  BASELINE.json's north_star
// asks for "pinned-host staging so the existing synchronous Mat API is unchanged": three mechanisms, cheapest first.
//   1. rcv_pinned_alloc[_on]: page-locked storage owned by the library.  The _on form places the pages on the NUMA
//      node of the GPU that will DMA them (mmap + mbind + cudaHostRegister) when the box exposes more than one node.
//   2. rcv_host_register / rcv_host_unregister: page-lock a caller-owned buffer in place (the Rust Mat wrapper
//      registers when `ensure_size` (re)allocates and unregisters in Drop -- INTEGRATION.md).  Registered ranges
//      are remembered, so a plain RCV_HOST Mat inside one takes the pinned (banded, direct-DMA) pipeline.
//      "host.auto_register" = 1 does the same on first sight of a buffer; OFF by default because a buffer freed
//      behind the library's back leaves a stale registration (the caller must then call rcv_host_unregister).
//   3. everything else (unknown pageable memory): abi.cu's bounce pipeline -- this file's memcpy pool copies
//      row bands between the Mat and per-slot pinned bounce buffers while the copy engines move the previous band.
#include "rcv_internal.cuh"

#include <dirent.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace rcv {

// ---------------------------------------------------------------------------------------------------------
// topology (sysfs)
// ---------------------------------------------------------------------------------------------------------
static bool gpu_sysfs_dir(int device, char *out, size_t cap) {
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  for (char *p = bus; *p; ++p)
    if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
  snprintf(out, cap, "/sys/bus/pci/devices/%s", bus);
  return true;
}

static bool read_small_file(const char *path, char *buf, size_t cap) {
  FILE *f = fopen(path, "r");
  if (!f) return false;
  size_t n = fread(buf, 1, cap - 1, f);
  fclose(f);
  buf[n] = 0;
  return n > 0;
}

static int online_numa_nodes() {
  DIR *d = opendir("/sys/devices/system/node");
  if (!d) return 1;
  int n = 0;
  while (dirent *e = readdir(d))
    if (strncmp(e->d_name, "node", 4) == 0 && e->d_name[4] >= '0' && e->d_name[4] <= '9') ++n;
  closedir(d);
  return n > 0 ? n : 1;
}

int gpu_numa_node(int device) {
  char dir[96], path[128], buf[32];
  if (!gpu_sysfs_dir(device, dir, sizeof(dir))) return -1;
  snprintf(path, sizeof(path), "%s/numa_node", dir);
  if (!read_small_file(path, buf, sizeof(buf))) return -1;
  return atoi(buf);
}

// "0-15,32-47" -> cpu_set_t
static bool parse_cpulist(const char *s, cpu_set_t *set) {
  CPU_ZERO(set);
  bool any = false;
  while (*s) {
    char *end;
    long a = strtol(s, &end, 10);
    if (end == s) break;
    long b = a;
    if (*end == '-') {
      s = end + 1;
      b = strtol(s, &end, 10);
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
      CPU_SET((int)c, set);
      any = true;
    }
    s = end;
    while (*s == ',' || *s == '\n' || *s == ' ') ++s;
  }
  return any;
}

void bind_thread_to_gpu(int device) {
  char dir[96], path[128], buf[1024];
  if (!gpu_sysfs_dir(device, dir, sizeof(dir))) return;
  snprintf(path, sizeof(path), "%s/local_cpulist", dir);
  if (!read_small_file(path, buf, sizeof(buf))) return;
  cpu_set_t local, allowed, both;
  if (!parse_cpulist(buf, &local)) return;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return;
  CPU_AND(&both, &local, &allowed);
  if (CPU_COUNT(&both) > 0 && CPU_COUNT(&both) < CPU_COUNT(&allowed)) sched_setaffinity(0, sizeof(both), &both);
}

// ---------------------------------------------------------------------------------------------------------
// registry of page-locked host ranges
// ---------------------------------------------------------------------------------------------------------
namespace {
enum PinKind { PIN_CUDA_ALLOC, PIN_MMAP, PIN_USER, PIN_AUTO };
struct PinRec {
  size_t bytes;
  PinKind kind;
  void *reg_base;    // what cudaHostRegister was given (page-rounded for caller-owned buffers)
  size_t map_bytes;  // PIN_MMAP: length of the mapping
  uint64_t stamp;
};
std::mutex g_pin_mu;
std::map<uintptr_t, PinRec> g_pins;  // keyed by user pointer
std::atomic<size_t> g_pin_count{0};
uint64_t g_pin_clock = 0;
size_t g_auto_bytes = 0;

// the record whose range contains [p, p+bytes); g_pin_mu held
std::map<uintptr_t, PinRec>::iterator find_range(uintptr_t p, size_t bytes) {
  auto it = g_pins.upper_bound(p);
  if (it == g_pins.begin()) return g_pins.end();
  --it;
  if (p >= it->first && p + bytes <= it->first + it->second.bytes) return it;
  return g_pins.end();
}

int register_range(void *ptr, size_t bytes, PinKind kind) {
  // exact range first; some drivers want page granularity: retry on the enclosing pages (they belong to the
  // same mapping -- a large Vec<u8> is its own mmap, a small one sits inside the heap's)
  void *base = ptr;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    const uintptr_t pg = (uintptr_t)sysconf(_SC_PAGESIZE);
    const uintptr_t lo = (uintptr_t)ptr & ~(pg - 1), hi = ((uintptr_t)ptr + bytes + pg - 1) & ~(pg - 1);
    base = (void *)lo;
    e = cudaHostRegister(base, hi - lo, cudaHostRegisterPortable);
  }
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return fail(RCV_ERR_ARG, "host range %p+%zu overlaps memory that is already page-locked", ptr, bytes);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
  g_pins[(uintptr_t)ptr] = PinRec{bytes, kind, base, 0, ++g_pin_clock};
  g_pin_count.store(g_pins.size());
  if (kind == PIN_AUTO) g_auto_bytes += bytes;
  return RCV_OK;
}

void release(std::map<uintptr_t, PinRec>::iterator it) {
  const PinRec r = it->second;
  void *user = (void *)it->first;
  g_pins.erase(it);
  g_pin_count.store(g_pins.size());
  switch (r.kind) {
    case PIN_CUDA_ALLOC: cudaFreeHost(user); break;
    case PIN_MMAP:
      cudaHostUnregister(r.reg_base);
      munmap(user, r.map_bytes);
      break;
    case PIN_AUTO: g_auto_bytes -= r.bytes;  // fall through
    case PIN_USER: cudaHostUnregister(r.reg_base); break;
  }
  cudaGetLastError();
}
}  // namespace

bool host_range_pinned(const void *p, size_t bytes) {
  if (g_pin_count.load(std::memory_order_relaxed) == 0) return false;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  auto it = find_range((uintptr_t)p, bytes);
  if (it == g_pins.end()) return false;
  it->second.stamp = ++g_pin_clock;
  return true;
}

bool host_auto_register(const void *p, size_t bytes) {
  if (bytes < (size_t)opt_get("host.auto_register_min_bytes", 1 << 20)) return false;
  const size_t budget = (size_t)opt_get("host.auto_register_max_bytes", (int64_t)1 << 31);
  std::lock_guard<std::mutex> lk(g_pin_mu);
  if (find_range((uintptr_t)p, bytes) != g_pins.end()) return true;
  // a buffer that overlaps a cached auto range was reallocated: the old registration is stale
  for (auto it = g_pins.begin(); it != g_pins.end();) {
    auto cur = it++;
    if (cur->second.kind == PIN_AUTO && cur->first < (uintptr_t)p + bytes && (uintptr_t)p < cur->first + cur->second.bytes)
      release(cur);
  }
  while (g_auto_bytes + bytes > budget) {  // LRU eviction
    auto victim = g_pins.end();
    for (auto it = g_pins.begin(); it != g_pins.end(); ++it)
      if (it->second.kind == PIN_AUTO && (victim == g_pins.end() || it->second.stamp < victim->second.stamp)) victim = it;
    if (victim == g_pins.end()) return false;
    release(victim);
  }
  return register_range(const_cast<void *>(p), bytes, PIN_AUTO) == RCV_OK;
}

void host_registry_shutdown() {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (auto it = g_pins.begin(); it != g_pins.end();) {
    auto cur = it++;
    if (cur->second.kind == PIN_AUTO || cur->second.kind == PIN_USER) release(cur);  // library-owned storage stays valid
  }
}

// ---------------------------------------------------------------------------------------------------------
// the memcpy pool
// ---------------------------------------------------------------------------------------------------------
namespace {
struct CopyTask {
  uint8_t *d;
  const uint8_t *s;
  size_t dstep, sstep, rb;
  int rows;
  std::atomic<int> *pending;
};

void run_task(const CopyTask &t) {
  if (t.dstep == t.rb && t.sstep == t.rb) {
    memcpy(t.d, t.s, t.rb * (size_t)t.rows);
  } else {
    for (int r = 0; r < t.rows; ++r) memcpy(t.d + (size_t)r * t.dstep, t.s + (size_t)r * t.sstep, t.rb);
  }
  t.pending->fetch_sub(1, std::memory_order_acq_rel);
}

struct CopyPool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<CopyTask> q;
  bool stop = false;

  void start(int n) {
    for (int i = 0; i < n; ++i)
      threads.emplace_back([this] {
        for (;;) {
          CopyTask t;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [this] { return stop || !q.empty(); });
            if (q.empty()) return;  // stop requested and nothing left
            t = q.front();
            q.pop_front();
          }
          run_task(t);
        }
      });
  }
  bool try_pop(CopyTask *t) {
    std::lock_guard<std::mutex> lk(mu);
    if (q.empty()) return false;
    *t = q.front();
    q.pop_front();
    return true;
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto &t : threads) t.join();
  }
};

CopyPool *g_pool = nullptr;
std::once_flag g_pool_once;

CopyPool *pool() {
  std::call_once(g_pool_once, [] {
    cpu_set_t allowed;
    int cpus = 4;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) cpus = CPU_COUNT(&allowed);
    int n = cpus / 4;
    if (n < 2) n = 2;
    if (n > 8) n = 8;
    n = (int)opt_get("host.copy_threads", n);
    g_pool = new CopyPool();
    if (n > 0) g_pool->start(n);
  });
  return g_pool;
}
}  // namespace

void host_copy2d(void *dst, size_t dstep, const void *src, size_t sstep, size_t row_bytes, int rows) {
  if (rows <= 0 || row_bytes == 0) return;
  const size_t total = row_bytes * (size_t)rows;
  if (total < ((size_t)256 << 10)) {  // not worth a hand-off
    std::atomic<int> one{1};
    run_task(CopyTask{(uint8_t *)dst, (const uint8_t *)src, dstep, sstep, row_bytes, rows, &one});
    return;
  }
  CopyPool *cp = pool();
  const size_t chunk = (size_t)opt_get("host.copy_chunk_bytes", 512 << 10);
  int rows_per = (int)(chunk / row_bytes);
  if (rows_per < 1) rows_per = 1;
  const int ntasks = (rows + rows_per - 1) / rows_per;
  std::atomic<int> pending{ntasks};
  {
    std::lock_guard<std::mutex> lk(cp->mu);
    for (int i = 0; i < ntasks; ++i) {
      const int r0 = i * rows_per, r1 = r0 + rows_per < rows ? r0 + rows_per : rows;
      cp->q.push_back(CopyTask{(uint8_t *)dst + (size_t)r0 * dstep, (const uint8_t *)src + (size_t)r0 * sstep, dstep, sstep,
                               row_bytes, r1 - r0, &pending});
    }
  }
  cp->cv.notify_all();
  // the caller works too (tasks of other callers included: everything in the queue is short), then waits
  CopyTask t;
  while (pending.load(std::memory_order_acquire) > 0) {
    if (cp->try_pop(&t))
      run_task(t);
    else
      sched_yield();
  }
}

// ---------------------------------------------------------------------------------------------------------
// drain threads: one per GPU, "wait for the band's D2H event, then copy the band from the bounce buffer to the Mat"
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Drainer {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<DrainJob> q;
  bool stop = false;
  void loop(int device) {
    cudaSetDevice(device);
    bind_thread_to_gpu(device);
    for (;;) {
      DrainJob j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return stop || !q.empty(); });
        if (q.empty()) return;
        j = q.front();
        q.pop_front();
      }
      if (cudaEventSynchronize(j.ev) != cudaSuccess) {
        cudaGetLastError();
        j.failed->store(1);
      } else {
        host_copy2d(j.dst, j.dstep, j.src, j.sstep, j.row_bytes, j.rows);
      }
      j.pending->fetch_sub(1, std::memory_order_acq_rel);
    }
  }
};
Drainer *g_drainers[16] = {};
std::mutex g_drainers_mu;
}  // namespace

void drain_submit(int device, const DrainJob &job) {
  Drainer *d;
  {
    std::lock_guard<std::mutex> lk(g_drainers_mu);
    d = g_drainers[device & 15];
    if (!d) {
      d = new Drainer();
      d->th = std::thread([d, device] { d->loop(device); });
      g_drainers[device & 15] = d;
    }
  }
  {
    std::lock_guard<std::mutex> lk(d->mu);
    d->q.push_back(job);
  }
  d->cv.notify_one();
}

void drain_shutdown() {
  std::lock_guard<std::mutex> lk(g_drainers_mu);
  for (int i = 0; i < 16; ++i) {
    Drainer *d = g_drainers[i];
    if (!d) continue;
    {
      std::lock_guard<std::mutex> lk2(d->mu);
      d->stop = true;
    }
    d->cv.notify_all();
    d->th.join();
    delete d;
    g_drainers[i] = nullptr;
  }
}

}  // namespace rcv

using namespace rcv;

// ---------------------------------------------------------------------------------------------------------
// ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int rcv_pinned_alloc_on(int32_t device, void **ptr, size_t bytes) {
  if (!ptr) return fail(RCV_ERR_ARG, "ptr is NULL");
  Ctx *c = device < 0 ? ctx_default() : ctx_get(device);
  if (!c) return RCV_ERR_NOT_INIT;
  if (bytes == 0) bytes = 1;
  const int node = c->numa_node;
  if (node >= 0 && online_numa_nodes() > 1 && opt_get("host.numa_place", 1) != 0) {
    const size_t pg = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = (bytes + pg - 1) / pg * pg;
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p != MAP_FAILED) {
      // MPOL_PREFERRED (1): pages come from the GPU's node when the cpuset allows it, from elsewhere otherwise
      unsigned long mask[16] = {};
      if (node < (int)(sizeof(mask) * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_mbind, p, len, 1 /*MPOL_PREFERRED*/, mask, sizeof(mask) * 8, 0UL);
      }
      madvise(p, len, MADV_HUGEPAGE);
      // first touch: the policy above decides the node, not the toucher's CPU
      for (size_t off = 0; off < len; off += pg) ((volatile uint8_t *)p)[off] = 0;
      cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
      if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pins[(uintptr_t)p] = PinRec{len, PIN_MMAP, p, len, ++g_pin_clock};
        g_pin_count.store(g_pins.size());
        *ptr = p;
        return RCV_OK;
      }
      cudaGetLastError();
      munmap(p, len);
    }
  }
  RCV_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pins[(uintptr_t)*ptr] = PinRec{bytes, PIN_CUDA_ALLOC, *ptr, 0, ++g_pin_clock};
  g_pin_count.store(g_pins.size());
  return RCV_OK;
}

int rcv_pinned_alloc(void **ptr, size_t bytes) { return rcv_pinned_alloc_on(-1, ptr, bytes); }

int rcv_pinned_free(void *ptr) {
  if (!ptr) return RCV_OK;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  auto it = g_pins.find((uintptr_t)ptr);
  if (it == g_pins.end() || (it->second.kind != PIN_CUDA_ALLOC && it->second.kind != PIN_MMAP))
    return fail(RCV_ERR_ARG, "%p was not returned by rcv_pinned_alloc", ptr);
  release(it);
  return RCV_OK;
}

int rcv_host_register(void *ptr, size_t bytes) {
  if (!ptr || bytes == 0) return fail(RCV_ERR_ARG, "rcv_host_register: NULL or empty range");
  if (!ctx_default()) return RCV_ERR_NOT_INIT;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  auto it = find_range((uintptr_t)ptr, bytes);
  if (it != g_pins.end()) {
    if (it->second.kind == PIN_AUTO) {  // promote: the caller now owns the lifetime
      g_auto_bytes -= it->second.bytes;
      it->second.kind = PIN_USER;
    }
    return RCV_OK;
  }
  return register_range(ptr, bytes, PIN_USER);
}

int rcv_host_unregister(void *ptr) {
  if (!ptr) return RCV_OK;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  auto it = g_pins.find((uintptr_t)ptr);
  if (it == g_pins.end()) return RCV_OK;  // never registered (or evicted): nothing to undo
  if (it->second.kind != PIN_USER && it->second.kind != PIN_AUTO)
    return fail(RCV_ERR_ARG, "%p is library-owned pinned storage: free it with rcv_pinned_free", ptr);
  release(it);
  return RCV_OK;
}

}  // extern "C"
