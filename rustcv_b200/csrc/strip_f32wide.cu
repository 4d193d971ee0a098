// strip_f32wide.cu -- separable 9x9 .. 15x15 filters on f32 (gray and BGR) in the TMA strip pipeline.
//   SepF32WideOp<KS, CN>: GaussianBlur / sepFilter2D with 9..15 taps; fmaf chains in the oracle's order
//   (orc_sepfilter_f32: ascending taps, row pass before column pass) -- bit-identical.
// Shape: like GaussQ8WideOp (strip_gaussq8_wide.cuh) -- 16-row chunks (2 * HV <= 14 warm-up rows, border patches one
// chunk back), 8 warps per CTA, rows in pairs with the window of row-filtered rows shifting two slots per pair; the
// neighbours of a row come by shuffle from up to 6 lanes away (15 taps on BGR reach 21 floats), so the op takes
// HALO_LANES = ceil(P * CN / 4) halo lanes per side.
#include "strip_f32_gather.cuh"

namespace rcv {

template <int KS, int CN>
struct SepF32WideOp {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = 4 * CN;
  static constexpr int REACH = P * CN;
  static constexpr int HALO_LANES = (REACH + 3) / 4;
  static constexpr int NOUT = 1;
  static constexpr int UNROLL = 2;
  static constexpr bool SINGLE_PATH = true;
  static constexpr int BAND_ROWS = 7 * 16 - 2 * HV;
  static_assert(KS >= 9 && KS <= 15 && (KS & 1), "kernel size");
  float win[KS][4];  // slots 0..KS-2: previous row-filtered rows, oldest first (at a pair's first row); slot KS-1: that first row
  float kx[KS], ky[KS];

  __device__ __forceinline__ void init(const StripParams &p) {
#pragma unroll
    for (int i = 0; i < KS; ++i) {
      kx[i] = p.ftaps[i];
      ky[i] = p.ftaps[KS + i];
    }
  }
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < KS; ++j)
#pragma unroll
      for (int h = 0; h < 4; ++h) win[j][h] = 0.0f;
  }
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &) {}  // SINGLE_PATH: never called

  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    constexpr int U = J8 & 1;
    float x[4 + 2 * REACH], h[4];
    gather_row<REACH>(q, x);  // shuffles: executed by the whole warp whether or not the row emits
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < KS; ++j) acc = fmaf(kx[j], x[REACH + c + (j - P) * CN], acc);
      h[c] = acc;
    }
    float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (emit) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < KS - 1; ++i) acc = fmaf(ky[i], win[i + U][c], acc);  // oldest row first
        v[c] = fmaf(ky[KS - 1], h[c], acc);
      }
    }
    if (U == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) win[KS - 1][c] = h[c];
    } else {
#pragma unroll
      for (int i = 0; i < KS - 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) win[i][c] = win[i + 2][c];
#pragma unroll
      for (int c = 0; c < 4; ++c) win[KS - 2][c] = h[c];
    }
    if (!emit) return;
    float *o = (float *)outp[0];
    if (nvalid == 16 && vec) {
      *(float4 *)o = make_float4(v[0], v[1], v[2], v[3]);
    } else if (nvalid > 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c * 4 < nvalid) o[c] = v[c];
    }
  }
};

template <int KS>
static int launch_sepf32wide_ks(Ctx *c, const DBatch &src, const DBatch &dst, const float *taps, cudaStream_t s) {
  // gray: under 170 registers per thread, so 12 warps fit (2 stages of 16 rows per warp = 192 KB per CTA)
  if (src.v.cn == 1) return launch_strip<SepF32WideOp<KS, 1>, 2, 12, 16>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
  if (src.v.cn == 3) {
    if constexpr (KS <= 9) return launch_strip<SepF32WideOp<KS, 3>, 2, 12, 16>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
    else return launch_strip<SepF32WideOp<KS, 3>, kS, 8, 16>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
  }
  return RCV_ERR_UNSUPPORTED;
}

// separable f32, gray or BGR, kw == kh in {9, 11, 13, 15}; RCV_ERR_UNSUPPORTED otherwise
int launch_sepf32wide_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky, int kh,
                            cudaStream_t s) {
  if (!strip_path_ok(src, 16, 32) || src.v.depth != RCV_F32 || kw != kh || kw < 9 || kw > 15 || !(kw & 1)) return RCV_ERR_UNSUPPORTED;
  if (opt_get("sepf32.no_wide", 0) != 0) return RCV_ERR_UNSUPPORTED;
  float taps[30];
  for (int i = 0; i < kw; ++i) {
    taps[i] = kx[i];
    taps[kw + i] = ky[i];
  }
  switch (kw) {
    case 9: return launch_sepf32wide_ks<9>(c, src, dst, taps, s);
    case 11: return launch_sepf32wide_ks<11>(c, src, dst, taps, s);
    case 13: return launch_sepf32wide_ks<13>(c, src, dst, taps, s);
    default: return launch_sepf32wide_ks<15>(c, src, dst, taps, s);
  }
}

}  // namespace rcv
