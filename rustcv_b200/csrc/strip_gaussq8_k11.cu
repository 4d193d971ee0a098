// strip_gaussq8_k11.cu -- 11x11 any-sigma Gaussians on u8.
// Where every tap lies within the adjacent lanes (CN * 5 <= 16 bytes) the horizontal-first op of strip_gaussq8.cuh
// serves 11 taps too: 8 warps per CTA (its 10 partial sums x 16 samples need up to 255 registers), 16-row chunks.
// 532 -> 345 instructions per row on BGR against GaussQ8WideOp (two horizontal passes, a shifting window), which keeps
// the channel counts that reach further (4 channels) and is selectable with the option "gauss.wide_windowed".
#include "strip_gaussq8.cuh"
#include "strip_gaussq8_wide.cuh"

namespace rcv {

int launch_gaussq8_k11(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, const int32_t *ky, cudaStream_t s) {
  if (opt_get("gauss.wide_windowed", 0) != 0) return launch_gaussq8_wide_ks<11>(c, src, dst, kx, ky, s);
  int32_t tx[8] = {0}, ty[8] = {0};
  for (int i = 0; i <= 11 / 2; ++i) {
    tx[i] = kx[i];
    ty[i] = ky[i];
  }
  switch (src.v.cn) {
    case 1: return launch_strip<GaussQ8Op<1, 11>, kS, 8, 16>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty, nullptr, 0, true);
    case 2: return launch_strip<GaussQ8Op<2, 11>, kS, 8, 16>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty, nullptr, 0, true);
    case 3: return launch_strip<GaussQ8Op<3, 11>, kS, 8, 16>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty, nullptr, 0, true);
  }
  return launch_gaussq8_wide_ks<11>(c, src, dst, kx, ky, s);
}

}  // namespace rcv
