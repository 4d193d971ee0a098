// strip_f2d_u8.cu -- dense 3x3 / 5x5 / 7x7 filter2D on u8 (1..4 channels) in the TMA strip pipeline.
//   acc = delta; acc = fmaf(k[i][j], (float)p[y+i-P][x+(j-P)*CN], acc) in row-major tap order;
//   out = saturate_u8(rint(acc))           (oracle: orc_filter2d_u8; rint = round half to even)
// Transposed form: a lane keeps, for each of its 16 samples, the running sums of the KS-1 output rows that have
// started but not finished.  A new source row is the ky-th row of output row r + P - ky: it extends every pending
// sum by its KS taps of that kernel row (in kx order) and starts a new one from delta, so each output sees its
// taps in exactly the oracle's row-major order -- bit-identical f32 -- while the state is 16 floats per pending row
// (64 registers at 5x5) instead of a window of KS-1 widened source rows WITH their halo columns (4 x 28 floats at
// 5x5 BGR: the windowed 5x5 needed 255 registers and measured below the general kernel).  Bytes become floats once
// per source row with the 2^23 trick (PRMT into 0x4B0000bb, minus 8388608.0f: exact); the taps are read from the
// kernel parameters (constant bank operands of the FFMAs), not held in registers; the result is rounded by adding
// 8388608.0f after clamping to [0, 255] (the FADD rounds to nearest even, like rint).
#include "strip_pipeline.cuh"

namespace rcv {

template <int CN, int KS>
struct Filter2dU8Op {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = CN;
  static constexpr int NOUT = 1;
  static constexpr int UNROLL = 2;
  static constexpr bool HOIST_WARM = true;
  static constexpr int HB = P * CN;          // halo bytes on each side of the lane's 16
  static constexpr int XW = 16 + 2 * HB;     // element columns -HB .. 15+HB
  static constexpr int NS = KS - 1;          // pending output rows
  static constexpr int HW = (HB + 3) / 4;    // halo words on each side
  static_assert((KS == 3 || KS == 5 || KS == 7) && HB <= 12, "3x3 / 5x5 / 7x7, halo within three words");
  float acc[NS][16];
  const StripParams *prm;  // taps: ftaps[ky * KS + kx], then delta

  __device__ __forceinline__ void init(const StripParams &p) { prm = &p; }
  __device__ __forceinline__ void reset() {}  // the 2*P warm-up rows start every pending sum
  __device__ __forceinline__ float tap(int ky, int kx) const { return prm->ftaps[ky * KS + kx]; }

  static __device__ __forceinline__ float byte_to_float(uint32_t w, int b) {
    return __uint2float_rn((w >> (8 * b)) & 0xFFu);  // exact; ptxas: I2F.U8 with a byte selector (conversion pipe)
  }
  __device__ __forceinline__ void widen(const uint4 &q, float (&x)[XW]) const {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t wl[HW], wr[HW];  // the left lane's last HW words, the right lane's first HW words
#pragma unroll
    for (int i = 0; i < HW; ++i) {
      wl[i] = __shfl_up_sync(0xffffffffu, w[4 - HW + i], 1);
      wr[i] = __shfl_down_sync(0xffffffffu, w[i], 1);
    }
#pragma unroll
    for (int e = 0; e < HB; ++e) {
      const int bl = 4 * HW - HB + e;  // byte of the left lane's last HW words
      x[e] = byte_to_float(wl[bl >> 2], bl & 3);
      x[HB + 16 + e] = byte_to_float(wr[e >> 2], e & 3);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) x[HB + e] = byte_to_float(w[e >> 2], e & 3);
  }
  // one kernel row onto a running sum: s = fmaf(k[ky][kx], x[e + (kx - P) * CN], s), kx ascending
  __device__ __forceinline__ float chain(float s, int ky, const float (&x)[XW], int e) const {
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) s = fmaf(tap(ky, kx), x[e + kx * CN], s);
    return s;
  }

  // Row J8 of a band (0 .. 2P-1) is kernel row ky of the outputs whose sums exist already: ky = 0 .. J8.
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    float x[XW];
    widen(q, x);
    constexpr int M = J8 < NS - 1 ? J8 : NS - 1;
    const float delta = prm->ftaps[KS * KS];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
#pragma unroll
      for (int m = M; m >= 0; --m)  // ascending state index: every update reads the old value of the next slot
        acc[NS - 1 - m][e] = chain(m == 0 ? delta : acc[NS - m][e], m, x, e);
    }
  }

  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    float x[XW];
    widen(q, x);
    const float delta = prm->ftaps[KS * KS];
    uint32_t r[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float o = chain(acc[0][e], KS - 1, x, e);  // the last kernel row completes output row r - P
#pragma unroll
      for (int i = 0; i < NS - 1; ++i) acc[i][e] = chain(acc[i + 1][e], KS - 2 - i, x, e);
      acc[NS - 1][e] = chain(delta, 0, x, e);
      // saturate_u8(rint(o)): one saturating conversion (round to nearest even; NaN -> 0 like fmaxf(NaN, 0))
      asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r[e]) : "f"(o));
    }
    if (!FAST && !emit) return;
    uint32_t ow[4];
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      const uint32_t lo = __byte_perm(r[4 * w4], r[4 * w4 + 1], 0x0040);
      const uint32_t hi = __byte_perm(r[4 * w4 + 2], r[4 * w4 + 3], 0x0040);
      ow[w4] = __byte_perm(lo, hi, 0x5410);
    }
    uint8_t *o = outp[0];
    if (FAST) {
      if (nvalid == 16) *(uint4 *)o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else if (nvalid == 16 && vec) {
      *(uint4 *)o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else if (nvalid > 0) {
#pragma unroll
      for (int b = 0; b < 16; ++b)
        if (b < nvalid) o[b] = (uint8_t)(ow[b >> 2] >> ((b & 3) * 8));
    }
  }
};

template <int KS>
static int launch_f2d_u8_ks(Ctx *c, const DBatch &src, const DBatch &dst, const float *taps, int ntaps, cudaStream_t s) {
  // 7x7 keeps 6 pending rows x 16 samples: 12 warps per CTA at <= 168 registers each
  constexpr int NW = KS == 7 ? 12 : kNW;
  switch (src.v.cn) {
    case 1: return launch_strip<Filter2dU8Op<1, KS>, kS, NW>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, ntaps);
    case 2: return launch_strip<Filter2dU8Op<2, KS>, kS, NW>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, ntaps);
    case 3: return launch_strip<Filter2dU8Op<3, KS>, kS, NW>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, ntaps);
    case 4: return launch_strip<Filter2dU8Op<4, KS>, kS, NW>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, ntaps);
  }
  return RCV_ERR_UNSUPPORTED;
}

int launch_filter2d_u8_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                             cudaStream_t s) {
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_U8 || kw != kh || (kw != 3 && kw != 5 && kw != 7)) return RCV_ERR_UNSUPPORTED;
  if (src.v.row_bytes() > (size_t)1 << 30) return RCV_ERR_UNSUPPORTED;
  float taps[50];
  for (int i = 0; i < kw * kh; ++i) taps[i] = k[i];
  taps[kw * kh] = delta;
  if (kw == 3) return launch_f2d_u8_ks<3>(c, src, dst, taps, 10, s);
  if (kw == 5) return launch_f2d_u8_ks<5>(c, src, dst, taps, 26, s);
  return launch_f2d_u8_ks<7>(c, src, dst, taps, 50, s);
}

}  // namespace rcv
