// strip_f32.cu -- f32 single-channel ops of the TMA strip pipeline (see strip_pipeline.cuh):
//   SepF32Op<KS>     separable KS x KS filter (GaussianBlur / sepFilter2D on gray f32), KS = 3, 5, 7
//   Filter2dF32Op<KS> dense KS x KS correlation, KS = 3, 5
// Operation order is the oracle's (orc_sepfilter_f32 / orc_filter2d_f32): fmaf chains in
// ascending tap order, row pass before column pass -- results are bit-identical.
#include "strip_pipeline.cuh"

namespace rcv {

template <int KS>
struct SepF32Op {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = 4;
  static constexpr int NOUT = 1;
  static constexpr int WIN = KS == 3 ? 2 : KS == 5 ? 4 : 8;
  static constexpr int UNROLL = WIN;
  static constexpr bool HOIST_WARM = KS <= 5;  // 2*HV == UNROLL: the window-filling rows run outside the steady loop
  float win[WIN][4];  // row-filtered previous rows
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
    for (int j = 0; j < WIN; ++j)
#pragma unroll
      for (int h = 0; h < 4; ++h) win[j][h] = 0.0f;
  }

  // row pass of one source row: h[c] = fmaf chain over kx, ascending, from 0
  __device__ __forceinline__ void rowpass(const uint4 &q, float (&h)[4]) const {
    float x[4 + 2 * P];  // columns -P .. 3+P
    x[P + 0] = __uint_as_float(q.x);
    x[P + 1] = __uint_as_float(q.y);
    x[P + 2] = __uint_as_float(q.z);
    x[P + 3] = __uint_as_float(q.w);
#pragma unroll
    for (int e = 0; e < P; ++e) {
      x[e] = __shfl_up_sync(0xffffffffu, x[P + 4 - P + e], 1);        // left lane's last P columns
      x[P + 4 + e] = __shfl_down_sync(0xffffffffu, x[P + e], 1);      // right lane's first P columns
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < KS; ++j) acc = fmaf(kx[j], x[c + j], acc);
      h[c] = acc;
    }
  }

  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    float h[4];
    rowpass(q, h);
#pragma unroll
    for (int c = 0; c < 4; ++c) win[J8 & (WIN - 1)][c] = h[c];
  }

  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    float h[4], v[4];
    rowpass(q, h);  // shuffles: executed by the whole warp whether or not the row emits
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = 0.0f;
#pragma unroll
      for (int i = 0; i < KS - 1; ++i) acc = fmaf(ky[i], win[(J8 + WIN * 8 - (KS - 1) + i) & (WIN - 1)][c], acc);
      v[c] = fmaf(ky[KS - 1], h[c], acc);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) win[J8 & (WIN - 1)][c] = h[c];
    if (!FAST && !emit) return;
    float *o = (float *)outp[0];
    if (FAST) {
      if (nvalid == 16) *(float4 *)o = make_float4(v[0], v[1], v[2], v[3]);
    } else if (nvalid == 16 && vec) {
      *(float4 *)o = make_float4(v[0], v[1], v[2], v[3]);
    } else if (nvalid > 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c * 4 < nvalid) o[c] = v[c];
    }
  }
};

int launch_sepf32cn_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky, int kh,
                          cudaStream_t s);
int launch_filter2d_f32cn_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                                cudaStream_t s);
int launch_sepf32wide_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky, int kh,
                            cudaStream_t s);

// separable f32, kw == kh in {3, 5, 7}: single channel here, 2..4 channels in strip_f32cn.cu; RCV_ERR_UNSUPPORTED otherwise
int launch_sepf32_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky, int kh,
                        cudaStream_t s) {
  if (kw == kh && kw >= 9) return launch_sepf32wide_strip(c, src, dst, kx, kw, ky, kh, s);  // 9..15 taps, gray / BGR
  if (src.v.cn > 1) return launch_sepf32cn_strip(c, src, dst, kx, kw, ky, kh, s);
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_F32 || src.v.cn != 1 || kw != kh) return RCV_ERR_UNSUPPORTED;
  float taps[14];
  if (kw != 3 && kw != 5 && kw != 7) return RCV_ERR_UNSUPPORTED;
  for (int i = 0; i < kw; ++i) {
    taps[i] = kx[i];
    taps[kw + i] = ky[i];
  }
  if (kw == 3) return launch_strip<SepF32Op<3>>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 6);
  if (kw == 5) return launch_strip<SepF32Op<5>>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 10);
  return launch_strip<SepF32Op<7>>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 14);
}

// dense f32: every channel count goes through the transposed-form op of strip_f32cn.cu (3x3 / 5x5 / 7x7)
int launch_filter2d_f32_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                              cudaStream_t s) {
  return launch_filter2d_f32cn_strip(c, src, dst, k, kw, kh, delta, s);
}

}  // namespace rcv
