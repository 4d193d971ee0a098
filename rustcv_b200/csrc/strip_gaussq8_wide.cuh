// strip_gaussq8_wide.cuh -- GaussQ8WideOp: 9x9 .. 15x15 Gaussians on u8 with arbitrary symmetric Q8 taps, in the TMA
// strip pipeline (instantiated per kernel size in strip_gaussq8_k{9,11,13,15}.cu).
//
// The arithmetic is GaussQ8Op's (strip_gaussq8.cuh): two samples per register as 16-bit lanes; vertical sum
// V <= 65280 exactly; V split into bytes Vh, Vl; A = sum kx Vh, B = sum kx Vl; out = (A + (B >> 8) + 128) >> 8
// == (sum ky kx p + 2^15) >> 16, the oracle's orc_sepfilter_u8_q8.  What differs is the shape around it:
//   * KS - 1 previous rows x 8 registers is up to 112 registers: the kernel runs 8 warps per CTA (a thread may then
//     hold 255 registers) and its chunks are 16 rows (the skeleton's border patches reach one chunk back, and a band
//     has to warm up 2 * HV <= 14 rows);
//   * the window does not rotate by compile-time indices (an unrolled period of 8 or 16 rows of ~500 instructions
//     each would not fit the instruction cache): rows are processed in pairs and the window shifts by two rows
//     per pair -- (KS - 2) * 8 register moves per two rows;
//   * tap j of a pixel lies (j - HV) * CN bytes away, up to 21 bytes = two lanes for 15 taps on BGR: neighbour words
//     come by shuffle from the lane that owns them, and the op takes HALO_LANES halo lanes per side.
#pragma once

#include "strip_pipeline.cuh"

namespace rcv {

template <int CN, int KS>
struct GaussQ8WideOp {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = CN;
  static constexpr int NOUT = 1;
  static constexpr int RB = P * CN;                  // bytes reached on each side
  static constexpr int EXT = (RB + 3) / 4;           // neighbour words on each side
  static constexpr int HALO_LANES = (RB + 15) / 16;
  static constexpr int UNROLL = 2;
  static constexpr bool SINGLE_PATH = true;          // one copy of the (long) row body
  static constexpr int BAND_ROWS = 7 * 16 - 2 * HV;  // 7 chunks of 16 fed rows per work item
  static_assert(KS >= 9 && KS <= 15 && (KS & 1), "kernel size");
  uint32_t win[KS][8];  // slots 0..KS-2: the previous rows, oldest first (at a pair's first row); slot KS-1: that first row
  uint32_t kx[HV + 1], ky[HV + 1];

  __device__ __forceinline__ void init(const StripParams &p) {
#pragma unroll
    for (int i = 0; i <= HV; ++i) {
      kx[i] = p.wtaps[i];
      ky[i] = p.wtaps[8 + i];
    }
  }
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < KS; ++j)
#pragma unroll
      for (int h = 0; h < 8; ++h) win[j][h] = 0;
  }

  __device__ __forceinline__ uint32_t sym(const uint32_t (&t)[KS], const uint32_t (&k)[HV + 1]) const {
    uint32_t acc = t[HV] * k[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) acc += (t[i] + t[KS - 1 - i]) * k[i];
    return acc;
  }

  // neighbour words by shuffle from the lane that owns them, odd-phase pairs by PRMT, then the KS taps of every pair
  __device__ __forceinline__ void hpass(const uint32_t (&X)[8], uint32_t (&out)[8]) const {
    constexpr int NWD = 4 + 2 * EXT;  // words -EXT .. 3+EXT at index +EXT
    uint32_t lo[NWD], hi[NWD], loS[NWD - 1], hiS[NWD - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo[k + EXT] = X[2 * k];
      hi[k + EXT] = X[2 * k + 1];
    }
#pragma unroll
    for (int o = 1; o <= EXT; ++o) {
      // word -o lives in lane - ceil(o / 4) at index (4 - o % 4) % 4; word 3 + o in lane + ceil(o / 4) at index (o - 1) % 4
      const int d = (o + 3) / 4;
      const int wl = (4 - o % 4) % 4, wr = (o - 1) % 4;
      lo[EXT - o] = __shfl_up_sync(0xffffffffu, X[2 * wl], d);
      hi[EXT - o] = __shfl_up_sync(0xffffffffu, X[2 * wl + 1], d);
      lo[EXT + 3 + o] = __shfl_down_sync(0xffffffffu, X[2 * wr], d);
      hi[EXT + 3 + o] = __shfl_down_sync(0xffffffffu, X[2 * wr + 1], d);
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
        out[2 * k + e] = sym(t, kx);
      }
  }

  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &) {}  // SINGLE_PATH: never called

  // J8 & 1 = position of the row in its pair
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    constexpr int U = J8 & 1;
    uint32_t in[8];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[2 * k] = __byte_perm(w[k], 0, 0x4240);
      in[2 * k + 1] = __byte_perm(w[k], 0, 0x4341);
    }
    uint32_t ow[4] = {0, 0, 0, 0};
    if (emit) {
      uint32_t Vh[8], Vl[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        uint32_t t[KS];
#pragma unroll
        for (int i = 0; i < KS - 1; ++i) t[i] = win[i + U][h];  // oldest first
        t[KS - 1] = in[h];
        const uint32_t V = sym(t, ky);  // <= 65280 per lane
        Vl[h] = V & 0x00FF00FFu;
        Vh[h] = (V >> 8) & 0x00FF00FFu;
      }
      uint32_t A[8], B[8];
      hpass(Vh, A);
      hpass(Vl, B);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t H0 = A[2 * k] + ((B[2 * k] >> 8) & 0x00FF00FFu) + 0x00800080u;
        const uint32_t H1 = A[2 * k + 1] + ((B[2 * k + 1] >> 8) & 0x00FF00FFu) + 0x00800080u;
        ow[k] = __byte_perm(H0, H1, 0x7351);
      }
    }
    // window: the pair's first row parks in slot KS-1; after the second row everything moves down two slots
    if (U == 0) {
#pragma unroll
      for (int h = 0; h < 8; ++h) win[KS - 1][h] = in[h];
    } else {
#pragma unroll
      for (int i = 0; i < KS - 2; ++i)
#pragma unroll
        for (int h = 0; h < 8; ++h) win[i][h] = win[i + 2][h];
#pragma unroll
      for (int h = 0; h < 8; ++h) win[KS - 2][h] = in[h];
    }
    if (!emit) return;
    uint8_t *o = outp[0];
    if (nvalid == 16 && vec) {
      *(uint4 *)o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else if (nvalid > 0) {
#pragma unroll
      for (int b = 0; b < 16; ++b)
        if (b < nvalid) o[b] = (uint8_t)(ow[b >> 2] >> ((b & 3) * 8));
    }
  }
};

// 8 warps per CTA, 16-row chunks: 3 stages x 8 KB per warp = 192 KB per CTA
template <int KS>
static inline int launch_gaussq8_wide_ks(Ctx *c, const DBatch &src, const DBatch &dst, const int32_t *kx, const int32_t *ky,
                                         cudaStream_t s) {
  int32_t tx[8] = {0}, ty[8] = {0};
  for (int i = 0; i <= KS / 2; ++i) {
    tx[i] = kx[i];
    ty[i] = ky[i];
  }
  switch (src.v.cn) {
    case 1: return launch_strip<GaussQ8WideOp<1, KS>, kS, 8, 16>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty, nullptr, 0, true);
    case 3: return launch_strip<GaussQ8WideOp<3, KS>, kS, 8, 16>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty, nullptr, 0, true);
    case 4: return launch_strip<GaussQ8WideOp<4, KS>, kS, 8, 16>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty, nullptr, 0, true);
  }
  return RCV_ERR_UNSUPPORTED;
}

}  // namespace rcv
