// The reciprocals the strip kernels decode work items with (rustcv_b200/csrc/fastdiv.h) against real division:
// every divisor up to 4096 and a spread of larger ones, n over [0, n_max) densely at both ends and around every
// multiple of d, for job sizes from one item to the 2^31 limit of the launcher.
#include <cstdio>
#include <cstdlib>

#include "fastdiv.h"

static int check(uint32_t d, uint64_t n_max) {
  uint32_t mul, sh;
  rcv::strip_fast_div(d, n_max, &mul, &sh);
  if (mul == 0) return d < 2 ? 0 : 1;  // 1: fell back to division (allowed, counted)
  auto q = [&](uint64_t n) { return (uint32_t)(((unsigned __int128)n * mul) >> 32) >> sh; };
  auto bad = [&](uint64_t n) {
    if (n >= n_max) return false;
    if (q(n) != (uint32_t)(n / d)) {
      std::printf("d=%u n_max=%llu n=%llu: %u != %llu\n", d, (unsigned long long)n_max, (unsigned long long)n, q(n),
                  (unsigned long long)(n / d));
      std::exit(1);
    }
    return false;
  };
  for (uint64_t n = 0; n < 5000 && n < n_max; ++n) bad(n);
  for (uint64_t n = n_max > 5000 ? n_max - 5000 : 0; n < n_max; ++n) bad(n);
  const uint64_t step = n_max / d / 997 + 1;
  for (uint64_t k = 0; k * d < n_max; k += step) {
    bad(k * d);
    bad(k * d + d - 1);
    if (k) bad(k * d - 1);
  }
  return 0;
}

int main() {
  const uint64_t sizes[] = {1, 2, 47, 2368, 44160, 1000003, (1ull << 24) + 5, (1ull << 31) - 1};
  long fallbacks = 0, cases = 0;
  for (uint64_t n_max : sizes) {
    for (uint32_t d = 1; d <= 4096; ++d, ++cases) fallbacks += check(d, n_max);
    for (uint32_t d = 4097; d < (1u << 30); d = d * 3 + 1, ++cases) fallbacks += check(d, n_max);
  }
  std::printf("fastdiv ok: %ld cases, %ld fell back to division\n", cases, fallbacks);
  return 0;
}
