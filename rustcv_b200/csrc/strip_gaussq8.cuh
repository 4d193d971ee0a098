// strip_gaussq8.cuh -- GaussQ8Op and its per-kernel-size launcher (instantiated in strip_gaussq8_k{3,5,7}.cu).
#pragma once

#include "strip_pipeline.cuh"

namespace rcv {

// ---------------------------------------------------------------------------------------
// Op: KS x KS Gaussian (KS = 3, 5, 7) on u8 with arbitrary symmetric Q8 taps (any sigma):
//   out = (sum_i sum_j ky_i kx_j p + 2^15) >> 16      (oracle: orc_sepfilter_u8_q8; == OpenCV)
// The sum is exact whatever the order, so the HORIZONTAL pass runs first, on the raw bytes, where two samples per
// register still work:  H = sum kx_j p <= 255 * 256 = 65280 fits a 16-bit lane (neighbour columns by shuffle,
// stride-CN taps by PRMT, as in Gauss5Op).  The vertical pass needs 24 bits; it runs on one sample per register,
// in transposed form: every sample keeps KS-1 partial sums s_0..s_{KS-2} (what the rows seen so far have
// contributed to the next KS-1 outputs) and a new row H costs KS multiply-adds per sample
//   out = s_0 + K[KS-1] H,   s_i = s_{i+1} + K[KS-2-i] H,   s_{KS-2} = K[0] H + 2^15
// with no window to rotate and no separate rounding step; the result is bits 16..23 of out.
// (Round 2, first version: vertical pass first in 16-bit lanes, then TWO horizontal passes -- shuffles and stride-CN
// gathers included -- over the high and low bytes of the vertical sums: 258 instructions per row at 5x5, now 185.)
// ---------------------------------------------------------------------------------------
template <int CN, int KS>
struct GaussQ8Op {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = CN;
  static constexpr int NOUT = 1;
  static constexpr int NS = KS - 1;  // partial sums per sample
  static constexpr int UNROLL = 2;   // no rotating window: the hot loop is as short as the instruction cache likes
  static constexpr bool HOIST_WARM = true;
  // warm-up rows cost a full horizontal pass here: taller bands than the default 5 chunks for the larger kernels
  // (measured, 5x5 on 32 x 4K BGR: 0.573 of the roofline at 36 rows, 0.585 at 60, 0.583 at 76, 0.573 at 116)
  static constexpr int BAND_ROWS = KS == 3 ? 5 * 8 - 2 * HV : KS <= 7 ? 8 * 8 - 2 * HV : 7 * 16 - 2 * HV;
  static constexpr int EXT = (HV * CN + 3) / 4;  // neighbour words needed on each side
  static constexpr int HALO_LANES = 1;           // every tap within the adjacent lanes (9 / 11 taps: CN * HV <= 16 bytes)
  static_assert(KS == 3 || KS == 5 || KS == 7 || KS == 9 || KS == 11, "kernel size");
  static_assert(EXT <= 4 && HV * CN <= 16, "taps beyond the adjacent lane");
  uint32_t sa[NS][8], sb[NS][8];  // partial sums of the low / high 16-bit lane's sample of register pair h
  uint32_t kx[HV + 1], ky[HV + 1];

  __device__ __forceinline__ void init(const StripParams &p) {
#pragma unroll
    for (int i = 0; i <= HV; ++i) {  // 9 / 11 taps arrive in the wide-tap slots (x: wtaps[0..7], y: wtaps[8..15])
      kx[i] = KS <= 7 ? p.taps_x[i & 3] : p.wtaps[i];
      ky[i] = KS <= 7 ? p.taps_y[i & 3] : p.wtaps[8 + i];
    }
  }
  __device__ __forceinline__ void reset() {}  // the 2*HV warm-up rows overwrite every partial sum

  // symmetric KS-tap sum of packed pairs: t[0..KS-1]
  __device__ __forceinline__ uint32_t sym(const uint32_t (&t)[KS], const uint32_t (&k)[HV + 1]) const {
    uint32_t acc = t[HV] * k[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) acc += (t[i] + t[KS - 1 - i]) * k[i];
    return acc;
  }

  // neighbour words by shuffle, odd-phase pairs by PRMT, then the KS taps of every output pair
  __device__ __forceinline__ void hpass(const uint4 &q, uint32_t (&out)[8]) const {
    constexpr int NWD = 4 + 2 * EXT;  // words -EXT .. 3+EXT at index +EXT
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t lo[NWD], hi[NWD], loS[NWD - 1], hiS[NWD - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo[k + EXT] = __byte_perm(w[k], 0, 0x4240);  // (b0, b2) as 16-bit lanes
      hi[k + EXT] = __byte_perm(w[k], 0, 0x4341);  // (b1, b3)
    }
#pragma unroll
    for (int e = 0; e < EXT; ++e) {
      // left lane's words 4-EXT+e -> our word -EXT+e ; right lane's word e -> our word 4+e
      lo[e] = __shfl_up_sync(0xffffffffu, lo[4 + e], 1);
      hi[e] = __shfl_up_sync(0xffffffffu, hi[4 + e], 1);
      lo[4 + EXT + e] = __shfl_down_sync(0xffffffffu, lo[EXT + e], 1);
      hi[4 + EXT + e] = __shfl_down_sync(0xffffffffu, hi[EXT + e], 1);
    }
#pragma unroll
    for (int i = 0; i < NWD - 1; ++i) {
      loS[i] = __byte_perm(lo[i], lo[i + 1], 0x5432);
      hiS[i] = __byte_perm(hi[i], hi[i + 1], 0x5432);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        uint32_t t[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          const int p = 4 * k + e + (j - HV) * CN + 4 * EXT;  // byte position relative to word -EXT
          const int wd = p >> 2, ph = p & 3;
          t[j] = ph == 0 ? lo[wd] : ph == 1 ? hi[wd] : ph == 2 ? loS[wd] : hiS[wd];
        }
        out[2 * k + e] = sym(t, kx);  // <= 65280 per lane
      }
  }

  // K[i], i = 0..KS-1: the full (symmetric) vertical tap array
  __device__ __forceinline__ uint32_t K(int i) const { return ky[i <= HV ? i : KS - 1 - i]; }

  // One row into the partial sums of one sample.  LIVE = how many sums (from the back) already hold something:
  // row j of a band has LIVE = j, so its warm-up costs j + 1 multiply-adds instead of KS.
  template <int LIVE, bool EMIT>
  __device__ __forceinline__ uint32_t vstep(uint32_t (&s)[NS][8], int h, uint32_t x) const {
    uint32_t out = 0;
    if (EMIT) out = s[0][h] + K(KS - 1) * x;
#pragma unroll
    for (int i = 0; i < NS - 1; ++i)
      if (NS - 1 - i <= LIVE) s[i][h] = s[i + 1][h] + K(KS - 2 - i) * x;
    s[NS - 1][h] = K(0) * x + 0x8000u;
    return out;
  }

  // The first 2*HV rows of a band only build the sums (J8 = the row of the band: the skeleton hoists them).
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    uint32_t H[8];
    hpass(q, H);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      vstep<(J8 < NS ? J8 : NS), false>(sa, h, H[h] & 0xFFFFu);
      vstep<(J8 < NS ? J8 : NS), false>(sb, h, H[h] >> 16);
    }
  }

  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    uint32_t H[8], ra[8], rb[8];
    hpass(q, H);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      ra[h] = vstep<NS, true>(sa, h, H[h] & 0xFFFFu);
      rb[h] = vstep<NS, true>(sb, h, H[h] >> 16);
    }
    if (!FAST && !emit) return;
    // pair 2k holds bytes (4k, 4k+2) of the row, pair 2k+1 bytes (4k+1, 4k+3); each result is byte 2 of its register
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t e02 = __byte_perm(ra[2 * k], rb[2 * k], 0x0062);          // (byte 4k, byte 4k+2, 0, 0)
      const uint32_t e13 = __byte_perm(ra[2 * k + 1], rb[2 * k + 1], 0x0062);  // (byte 4k+1, byte 4k+3, 0, 0)
      ow[k] = __byte_perm(e02, e13, 0x5140);
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
static inline int launch_gaussq8_ks(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, const int32_t *ky,
                             cudaStream_t s) {
  int32_t tx[4] = {0, 0, 0, 0}, ty[4] = {0, 0, 0, 0};
  for (int i = 0; i <= KS / 2; ++i) {
    tx[i] = kx[i];
    ty[i] = ky[i];
  }
  // 7x7 keeps 6 partial sums for each of 16 samples (96 registers): 12 warps per CTA at <= 168 registers each
  constexpr int NW = KS == 7 ? 12 : kNW;
  switch (src.v.cn) {
    case 1: return launch_strip<GaussQ8Op<1, KS>, kS, NW>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
    case 2: return launch_strip<GaussQ8Op<2, KS>, kS, NW>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
    case 3: return launch_strip<GaussQ8Op<3, KS>, kS, NW>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
    case 4: return launch_strip<GaussQ8Op<4, KS>, kS, NW>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
  }
  return RCV_ERR_UNSUPPORTED;
}



}  // namespace rcv
