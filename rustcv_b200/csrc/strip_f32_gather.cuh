// strip_f32_gather.cuh -- a lane's four floats of a tile row plus REACH floats on each side, the neighbours fetched
// with one shuffle each from the lane that owns them (shared by the multi-channel and the wide f32 strip ops).
#pragma once

#include "strip_pipeline.cuh"

namespace rcv {

// x[o + OFF] = float o of the row relative to this lane's first float, o = -REACH .. 3 + REACH
template <int REACH>
__device__ __forceinline__ void gather_row(const uint4 &q, float (&x)[4 + 2 * REACH]) {
  const float own[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
  for (int c = 0; c < 4; ++c) x[REACH + c] = own[c];
#pragma unroll
  for (int o = 1; o <= REACH; ++o) {
    // float -o lives in lane - ceil(o / 4) at index (4 - o % 4) % 4; float 3 + o in lane + ceil(o / 4) at index (o - 1) % 4
    const int d = (o + 3) / 4;
    x[REACH - o] = __shfl_up_sync(0xffffffffu, own[(4 - o % 4) % 4], d);
    x[REACH + 3 + o] = __shfl_down_sync(0xffffffffu, own[(o - 1) % 4], d);
  }
}

}  // namespace rcv
