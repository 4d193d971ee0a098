// strip_f32wide.cu -- separable 9x9 .. 15x15 filters on f32 (gray and BGR) in the TMA strip pipeline.
//   SepF32WideOp<KS, CN>: GaussianBlur / sepFilter2D with 9..15 taps; fmaf chains in the oracle's order
//   (orc_sepfilter_f32: ascending taps from 0.0f, row pass before column pass) -- bit-identical.
// Shape: 16-row chunks (2 * HV <= 14 warm-up rows, border patches one chunk back); the neighbours of a row come by
// shuffle from up to 6 lanes away (15 taps on BGR reach 21 floats), so the op takes HALO_LANES = ceil(P * CN / 4) halo
// lanes per side.  The column pass runs in TRANSPOSED form (second session of round 2, see Filter2dF32CnOp): a lane
// keeps the KS-1 pending column sums of its 4 floats; a row-filtered row h extends each (acc_i = fmaf(ky[KS-2-i],
// h, acc_{i+1})), completes the oldest and starts a new one (fmaf(ky[0], h, 0.0f)) -- the oracle's ascending order.
// Against the first version (a window of KS-1 row-filtered rows shifted two slots per row pair): no (KS-2) x 4
// register moves per pair, the 2 x KS taps are uniform-register operands instead of 30 registers, so BGR 11..15 taps
// fit 12 warps per CTA instead of 8, and the loop is unpredicated.
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
  static constexpr int BAND_ROWS = 7 * 16 - 2 * HV;
  static constexpr int NS = KS - 1;  // pending output rows
  static_assert(KS >= 9 && KS <= 15 && (KS & 1), "kernel size");  // 7 taps: measured slower than SepF32CnOp's 8-row window (0.79 vs 0.83)
  float acc[NS][4];
  const StripParams *prm;  // ftaps: kx[0..KS), ky[0..KS)

  __device__ __forceinline__ void init(const StripParams &p) { prm = &p; }
  __device__ __forceinline__ void reset() {}  // 2*HV warm-up rows rebuild every pending sum

  // row pass of this lane's 4 floats, then the column sums move up one slot; returns the completed row
  __device__ __forceinline__ void step(const uint4 &q, float (&v)[4]) {
    float x[4 + 2 * REACH];
    gather_row<REACH>(q, x);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float h = 0.0f;
#pragma unroll
      for (int j = 0; j < KS; ++j) h = fmaf(prm->ftaps[j], x[REACH + c + (j - P) * CN], h);
      v[c] = fmaf(prm->ftaps[KS + KS - 1], h, acc[0][c]);  // the newest row is the last tap of the oldest pending sum
#pragma unroll
      for (int i = 0; i < NS - 1; ++i) acc[i][c] = fmaf(prm->ftaps[KS + KS - 2 - i], h, acc[i + 1][c]);
      acc[NS - 1][c] = fmaf(prm->ftaps[KS], h, 0.0f);
    }
  }
  // warm-up rows run the full update: whatever the sums held before is pushed out within KS-1 rows
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    float v[4];
    step(q, v);
  }
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    float v[4];
    step(q, v);
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
static int launch_sepf32wide_ks(Ctx *c, const DBatch &src, const DBatch &dst, const float *taps, cudaStream_t s) {
  // 12 warps x 2 stages of 16 rows = 192 KB per CTA; every instance stays under 168 registers per thread
  if (src.v.cn == 1) return launch_strip<SepF32WideOp<KS, 1>, 2, 12, 16>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
  if (src.v.cn == 3) return launch_strip<SepF32WideOp<KS, 3>, 2, 12, 16>(c, src, &dst, 1, "sepf32.band_rows", s, nullptr, nullptr, taps, 2 * KS);
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
