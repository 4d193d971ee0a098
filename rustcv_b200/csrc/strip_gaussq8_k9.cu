// strip_gaussq8_k9.cu -- 9x9 instantiations of GaussQ8WideOp (CN = 1, 3, 4).
#include "strip_gaussq8_wide.cuh"

namespace rcv {

int launch_gaussq8_k9(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, const int32_t *ky, cudaStream_t s) {
  return launch_gaussq8_wide_ks<9>(c, src, dst, kx, ky, s);
}

}  // namespace rcv
