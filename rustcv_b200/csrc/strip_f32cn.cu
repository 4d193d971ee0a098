// strip_f32cn.cu -- multi-channel f32 ops of the TMA strip pipeline (see strip_pipeline.cuh):
//   SepF32CnOp<KS, CN>      separable KS x KS filter (GaussianBlur / sepFilter2D on f32 BGR / BGRA / 2-channel), KS = 3, 5, 7
//   Filter2dF32CnOp<KS, CN> dense KS x KS correlation, KS = 3, 5, 7, CN = 1..4 (transposed form: pending row sums)
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

// dense KS x KS correlation on CN interleaved channels (CN = 1: gray):
//   acc = delta; acc = fmaf(k[i][j], p[y+i-r][x+j-r], acc), row-major            (oracle: orc_filter2d_f32)
// Transposed form (second session of round 2; see Filter2dU8Op): a lane keeps, for each of its 4 floats, the running sums
// of the KS-1 output rows that have started but not finished.  A new source row -- gathered ONCE with its REACH
// neighbour floats per side, one shuffle each -- extends every pending sum by the KS taps of its kernel row and starts
// a new one from delta: every output accumulates in the oracle's row-major order (0 ULP), the state is 4 x (KS-1)
// registers instead of a window of KS-1 gathered rows (4 x 16 floats + 25 taps = 161 registers at 5x5 BGR), the taps
// are uniform-register operands of the FFMAs, and 7x7 fits (6 x 4 sums + 22 gathered floats at BGR).
template <int KS, int CN>
struct Filter2dF32CnOp {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = 4 * CN;
  static constexpr int REACH = P * CN;
  static constexpr int HALO_LANES = (REACH + 3) / 4;
  static constexpr int NOUT = 1;
  static constexpr int UNROLL = 2;
  static constexpr bool HOIST_WARM = true;
  static constexpr int XW = 4 + 2 * REACH;
  static constexpr int NS = KS - 1;  // pending output rows
  static_assert(KS == 3 || KS == 5 || KS == 7, "kernel size");
  float acc[NS][4];
  const StripParams *prm;  // taps: ftaps[ky * KS + kx], then delta

  __device__ __forceinline__ void init(const StripParams &p) { prm = &p; }
  __device__ __forceinline__ void reset() {}  // the 2*P warm-up rows start every pending sum
  // one kernel row onto a running sum, kx ascending
  __device__ __forceinline__ float chain(float s, int ky, const float (&x)[XW], int c) const {
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) s = fmaf(prm->ftaps[ky * KS + kx], x[c + kx * CN], s);
    return s;
  }
  // Row J8 of a band (0 .. 2P-1) is kernel row ky of the outputs whose sums exist already: ky = 0 .. J8.
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    float x[XW];
    gather_row<REACH>(q, x);
    constexpr int M = J8 < NS - 1 ? J8 : NS - 1;
    const float delta = prm->ftaps[KS * KS];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int m = M; m >= 0; --m)  // ascending state index: every update reads the old value of the next slot
        acc[NS - 1 - m][c] = chain(m == 0 ? delta : acc[NS - m][c], m, x, c);
    }
  }
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    float x[XW], v[4];
    gather_row<REACH>(q, x);
    const float delta = prm->ftaps[KS * KS];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = chain(acc[0][c], KS - 1, x, c);  // the last kernel row completes output row r - P
#pragma unroll
      for (int i = 0; i < NS - 1; ++i) acc[i][c] = chain(acc[i + 1][c], KS - 2 - i, x, c);
      acc[NS - 1][c] = chain(delta, 0, x, c);
    }
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

template <int KS>
static int launch_f2d_f32cn_ks(Ctx *c, const DBatch &src, const DBatch &dst, const float *taps, cudaStream_t s) {
  switch (src.v.cn) {
    case 1: return launch_strip<Filter2dF32CnOp<KS, 1>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, KS * KS + 1);
    case 2: return launch_strip<Filter2dF32CnOp<KS, 2>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, KS * KS + 1);
    case 3: return launch_strip<Filter2dF32CnOp<KS, 3>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, KS * KS + 1);
    case 4: return launch_strip<Filter2dF32CnOp<KS, 4>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, KS * KS + 1);
  }
  return RCV_ERR_UNSUPPORTED;
}

// dense f32, 1..4 channels, 3x3 / 5x5 / 7x7; RCV_ERR_UNSUPPORTED otherwise
int launch_filter2d_f32cn_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                                cudaStream_t s) {
  if (!strip_path_ok(src, 8, src.v.cn == 1 ? 8 : 16) || src.v.depth != RCV_F32 || src.v.cn < 1 || src.v.cn > 4 || kw != kh)
    return RCV_ERR_UNSUPPORTED;
  if (kw != 3 && kw != 5 && kw != 7) return RCV_ERR_UNSUPPORTED;
  float taps[50];
  for (int i = 0; i < kw * kh; ++i) taps[i] = k[i];
  taps[kw * kh] = delta;
  if (kw == 3) return launch_f2d_f32cn_ks<3>(c, src, dst, taps, s);
  if (kw == 5) return launch_f2d_f32cn_ks<5>(c, src, dst, taps, s);
  return launch_f2d_f32cn_ks<7>(c, src, dst, taps, s);
}

}  // namespace rcv
