// strip_gaussq8_k5.cu -- 5x5 instantiations of GaussQ8Op (CN = 1..4).
#include "strip_gaussq8.cuh"

namespace rcv {

int launch_gaussq8_k5(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, const int32_t *ky, cudaStream_t s) {
  return launch_gaussq8_ks<5>(c, src, dst, kx, ky, s);
}

}  // namespace rcv
