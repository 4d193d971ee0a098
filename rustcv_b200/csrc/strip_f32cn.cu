// strip_f32cn.cu -- multi-channel f32 ops of the TMA strip pipeline (see strip_pipeline.cuh):
//   SepF32CnOp<KS, CN>      separable KS x KS filter (GaussianBlur / sepFilter2D on f32 BGR / BGRA / 2-channel), KS = 3, 5, 7
//   Filter2dF32CnOp<KS, CN> dense KS x KS correlation, KS = 3, 5
// The single-channel ops (strip_f32.cu) take their row neighbours from the adjacent lane; with CN interleaved
// channels tap j of a pixel lies (j - P) * CN floats away, up to 9 (3 lanes) for a 7-tap filter on BGR, so these
// ops give up more than one halo lane per side (Op::HALO_LANES) and fetch each neighbour float with one shuffle
// from the lane that owns it.  Operation order is the oracle's (orc_sepfilter_f32 / orc_filter2d_f32): fmaf chains
// in ascending tap order, row pass before column pass -- bit-identical, like the single-channel ops.
#include "strip_f32_gather.cuh"

namespace rcv {

template <int KS, int CN>
struct SepF32CnOp {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = 4 * CN;
  static constexpr int REACH = P * CN;                  // floats needed on each side of the lane's four
  static constexpr int HALO_LANES = (REACH + 3) / 4;
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

  __device__ __forceinline__ void rowpass(const uint4 &q, float (&h)[4]) const {
    float x[4 + 2 * REACH];
    gather_row<REACH>(q, x);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < KS; ++j) acc = fmaf(kx[j], x[REACH + c + (j - P) * CN], acc);
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

// dense KS x KS correlation on CN interleaved channels: acc = delta; acc = fmaf(k[i][j], p[y+i-r][x+j-r], acc), row-major
template <int KS, int CN>
struct Filter2dF32CnOp {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = 4 * CN;
  static constexpr int REACH = P * CN;
  static constexpr int HALO_LANES = (REACH + 3) / 4;
  static constexpr int NOUT = 1;
  static constexpr int WIN = KS == 3 ? 2 : 4;
  static constexpr int UNROLL = WIN;
  static constexpr int XW = 4 + 2 * REACH;
  float win[WIN][XW];  // previous source rows with their neighbours
  float k[KS * KS];
  float delta;

  __device__ __forceinline__ void init(const StripParams &p) {
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) k[i] = p.ftaps[i];
    delta = p.ftaps[KS * KS];
  }
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < WIN; ++j)
#pragma unroll
      for (int h = 0; h < XW; ++h) win[j][h] = 0.0f;
  }
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    float x[XW];
    gather_row<REACH>(q, x);
#pragma unroll
    for (int c = 0; c < XW; ++c) win[J8 & (WIN - 1)][c] = x[c];
  }
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    float x[XW], v[4];
    gather_row<REACH>(q, x);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = delta;
#pragma unroll
      for (int i = 0; i < KS - 1; ++i)
#pragma unroll
        for (int j = 0; j < KS; ++j)
          acc = fmaf(k[i * KS + j], win[(J8 + WIN * 8 - (KS - 1) + i) & (WIN - 1)][REACH + c + (j - P) * CN], acc);
#pragma unroll
      for (int j = 0; j < KS; ++j) acc = fmaf(k[(KS - 1) * KS + j], x[REACH + c + (j - P) * CN], acc);
      v[c] = acc;
    }
#pragma unroll
    for (int c = 0; c < XW; ++c) win[J8 & (WIN - 1)][c] = x[c];
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

template <int KS>
static int launch_sepf32cn_ks(Ctx *c, const DBatch &src, const DBatch &dst, const float *taps, cudaStream_t s) {
  switch (src.v.cn) {
    case 2: return launch_strip<SepF32CnOp<KS, 2>>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
    case 3: return launch_strip<SepF32CnOp<KS, 3>>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
    case 4: return launch_strip<SepF32CnOp<KS, 4>>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
  }
  return RCV_ERR_UNSUPPORTED;
}

// separable f32, 2..4 channels, kw == kh in {3, 5, 7}; RCV_ERR_UNSUPPORTED otherwise
int launch_sepf32cn_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *kx, int kw, const float *ky, int kh,
                          cudaStream_t s) {
  if (!strip_path_ok(src, 8, 16) || src.v.depth != RCV_F32 || src.v.cn < 2 || src.v.cn > 4 || kw != kh) return RCV_ERR_UNSUPPORTED;
  if (kw != 3 && kw != 5 && kw != 7) return RCV_ERR_UNSUPPORTED;
  float taps[14];
  for (int i = 0; i < kw; ++i) {
    taps[i] = kx[i];
    taps[kw + i] = ky[i];
  }
  if (kw == 3) return launch_sepf32cn_ks<3>(c, src, dst, taps, s);
  if (kw == 5) return launch_sepf32cn_ks<5>(c, src, dst, taps, s);
  return launch_sepf32cn_ks<7>(c, src, dst, taps, s);
}

// dense f32: 3x3 on 3 or 4 channels, 5x5 on 3 channels (the 4-channel 5x5 window does not fit the register file)
int launch_filter2d_f32cn_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                                cudaStream_t s) {
  if (!strip_path_ok(src, 8, 16) || src.v.depth != RCV_F32 || (src.v.cn != 3 && src.v.cn != 4) || kw != kh) return RCV_ERR_UNSUPPORTED;
  if (kw == 5 && src.v.cn == 3) {  // 4 rows x 16 floats of window + 25 taps: 12 warps per CTA so that a thread may hold 168 registers
    float t5[26];
    for (int i = 0; i < 25; ++i) t5[i] = k[i];
    t5[25] = delta;
    return launch_strip<Filter2dF32CnOp<5, 3>, kS, 12>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, t5, 26);
  }
  if (kw != 3) return RCV_ERR_UNSUPPORTED;
  float taps[10];
  for (int i = 0; i < 9; ++i) taps[i] = k[i];
  taps[9] = delta;
  if (src.v.cn == 3) return launch_strip<Filter2dF32CnOp<3, 3>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, 10);
  return launch_strip<Filter2dF32CnOp<3, 4>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, 10);
}

}  // namespace rcv
