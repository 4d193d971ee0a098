// link_probe.cu -- what the host side of the box gives N GPUs at once (diagnostic, not product).
//
//   nvcc -O2 -std=c++17 -o scripts/_link_probe scripts/link_probe.cu -lpthread && scripts/_link_probe [MiB]
//
// One thread per GPU, pinned host buffers, plain cudaMemcpyAsync in both directions, timed on the host around a
// barrier-aligned burst (and per GPU with CUDA events).  Prints, for every GPU subset {0..n-1} and for some
// pairs: H2D alone, D2H alone, both at once -- per GPU and in total.  This is the ceiling bench.py's e2e figure
// is judged against at N > 1.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                 \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

struct Gpu {
  int dev;
  void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  cudaStream_t s1, s2;
  cudaEvent_t e0, e1, e2;
};

static std::string slurp(const std::string &path) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return "?";
  char buf[512];
  size_t n = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[n] = 0;
  while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
  return buf;
}

struct Barrier {
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    int g = gen.load();
    if (count.fetch_add(1) + 1 == n) {
      count.store(0);
      gen.fetch_add(1);
    } else {
      while (gen.load() == g) sched_yield();
    }
  }
};

// mode: 1 = H2D, 2 = D2H, 3 = both.  Returns wall seconds of the slowest GPU; per-GPU event ms in ms_out.
static double burst(std::vector<Gpu *> &gs, size_t bytes, int mode, int reps, std::vector<double> *gbs_each) {
  Barrier bar((int)gs.size());
  std::vector<double> secs(gs.size());
  std::vector<std::thread> th;
  for (size_t i = 0; i < gs.size(); ++i)
    th.emplace_back([&, i] {
      Gpu *g = gs[i];
      CK(cudaSetDevice(g->dev));
      double best = 1e30;
      for (int r = 0; r < reps; ++r) {
        CK(cudaDeviceSynchronize());
        bar.wait();
        auto t0 = std::chrono::steady_clock::now();
        if (mode & 1) CK(cudaMemcpyAsync(g->d_in, g->h_in, bytes, cudaMemcpyHostToDevice, g->s1));
        if (mode & 2) CK(cudaMemcpyAsync(g->h_out, g->d_out, bytes, cudaMemcpyDeviceToHost, g->s2));
        CK(cudaStreamSynchronize(g->s1));
        CK(cudaStreamSynchronize(g->s2));
        auto t1 = std::chrono::steady_clock::now();
        bar.wait();
        double s = std::chrono::duration<double>(t1 - t0).count();
        if (s < best) best = s;
      }
      secs[i] = best;
    });
  for (auto &t : th) t.join();
  double worst = 0;
  gbs_each->clear();
  for (size_t i = 0; i < gs.size(); ++i) {
    gbs_each->push_back(bytes / secs[i] / 1e9);
    if (secs[i] > worst) worst = secs[i];
  }
  return worst;
}

static void report(const char *label, std::vector<Gpu *> gs, size_t bytes) {
  static const char *names[4] = {"", "h2d", "d2h", "duplex"};
  printf("%-22s", label);
  for (int mode = 1; mode <= 3; ++mode) {
    std::vector<double> each;
    double worst = burst(gs, bytes, mode, 4, &each);
    double total = gs.size() * bytes / worst / 1e9;
    double mn = 1e30, mx = 0;
    for (double v : each) {
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
    printf("  %s: total %6.1f GB/s%s (per GPU %5.1f..%5.1f)", names[mode], total, mode == 3 ? " each way" : "", mn, mx);
  }
  printf("\n");
  fflush(stdout);
}

int main(int argc, char **argv) {
  size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 256;
  size_t bytes = mib << 20;
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  printf("GPUs: %d, %zu MiB per copy, cpus allowed: %ld, numa nodes: %s, Mems_allowed_list: ", n, mib,
         sysconf(_SC_NPROCESSORS_ONLN), slurp("/sys/devices/system/node/online").c_str());
  {
    FILE *f = fopen("/proc/self/status", "r");
    char line[256];
    while (f && fgets(line, sizeof(line), f))
      if (strncmp(line, "Mems_allowed_list", 17) == 0 || strncmp(line, "Cpus_allowed_list", 17) == 0) printf("%s ", strtok(line, "\n"));
    if (f) fclose(f);
    printf("\n");
  }
  std::vector<Gpu> gpus(n);
  for (int d = 0; d < n; ++d) {
    Gpu &g = gpus[d];
    g.dev = d;
    CK(cudaSetDevice(d));
    char bus[32];
    CK(cudaDeviceGetPCIBusId(bus, sizeof(bus), d));
    for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
    std::string dir = std::string("/sys/bus/pci/devices/") + bus;
    printf("  gpu%d %s numa_node=%s local_cpulist=%s link=%s x%s (max %s x%s)\n", d, bus, slurp(dir + "/numa_node").c_str(),
           slurp(dir + "/local_cpulist").c_str(), slurp(dir + "/current_link_speed").c_str(),
           slurp(dir + "/current_link_width").c_str(), slurp(dir + "/max_link_speed").c_str(),
           slurp(dir + "/max_link_width").c_str());
    CK(cudaHostAlloc(&g.h_in, bytes, cudaHostAllocPortable));
    CK(cudaHostAlloc(&g.h_out, bytes, cudaHostAllocPortable));
    memset(g.h_in, 1, bytes);
    memset(g.h_out, 2, bytes);
    CK(cudaMalloc(&g.d_in, bytes));
    CK(cudaMalloc(&g.d_out, bytes));
    CK(cudaStreamCreateWithFlags(&g.s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.s2, cudaStreamNonBlocking));
  }
  for (int k = 1; k <= n; k = k < 2 ? 2 : k * 2) {
    std::vector<Gpu *> gs;
    for (int d = 0; d < k; ++d) gs.push_back(&gpus[d]);
    char label[64];
    snprintf(label, sizeof(label), "gpus 0..%d", k - 1);
    report(label, gs, bytes);
  }
  for (int d = 1; d < n; ++d) {  // pairs: which GPUs share an uplink with gpu 0?
    char label[64];
    snprintf(label, sizeof(label), "pair {0,%d}", d);
    report(label, {&gpus[0], &gpus[d]}, bytes);
  }
  if (n >= 4) report("gpus {0,2,4,6}", n >= 8 ? std::vector<Gpu *>{&gpus[0], &gpus[2], &gpus[4], &gpus[6]} : std::vector<Gpu *>{&gpus[0], &gpus[2]}, bytes);
  // host memory itself: one memcpy thread per "GPU", same buffers
  for (int k = 1; k <= 8; k *= 2) {
    std::vector<std::thread> th;
    Barrier bar(k);
    std::vector<double> secs(k);
    for (int i = 0; i < k; ++i)
      th.emplace_back([&, i] {
        Gpu &g = gpus[i % n];
        bar.wait();
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < 3; ++r) memcpy(g.h_out, g.h_in, bytes);
        auto t1 = std::chrono::steady_clock::now();
        secs[i] = std::chrono::duration<double>(t1 - t0).count() / 3;
      });
    for (auto &t : th) t.join();
    double worst = 0;
    for (double s : secs) worst = s > worst ? s : worst;
    printf("host memcpy, %d thread(s): %.1f GB/s copied in total (read + write = 2x)\n", k, k * bytes / worst / 1e9);
  }
  return 0;
}
