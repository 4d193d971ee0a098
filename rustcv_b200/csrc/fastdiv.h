// fastdiv.h -- host-made reciprocals for the strip kernels' item decode (plain C++: also compiled by tests/cpp/fastdiv_test.cpp).
#pragma once

#include <stdint.h>

namespace rcv {

// n / d == umulhi(n, mul) >> sh for every n < n_max (Granlund-Montgomery round-up reciprocal, accepted only when its
// error bound covers n_max); mul = 0 tells the kernel to divide (d < 2, or no 32-bit multiplier is exact up to n_max).
static inline void strip_fast_div(uint32_t d, uint64_t n_max, uint32_t *mul, uint32_t *sh) {
  *mul = 0;
  *sh = 0;
  if (d < 2) return;
  for (uint32_t s = 0; s < 32; ++s) {
    const unsigned __int128 two = (unsigned __int128)1 << (32 + s);
    const unsigned __int128 m = (two + d - 1) / d;  // ceil(2^(32+s) / d)
    if (m >> 32) break;
    const unsigned __int128 e = m * d - two;        // 0 <= e < d; exact while n * e < 2^(32+s)
    if ((unsigned __int128)n_max * e < two) {
      *mul = (uint32_t)m;
      *sh = s;
      return;
    }
  }
}

}  // namespace rcv
