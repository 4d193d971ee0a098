// strip_gaussq8.cuh -- GaussQ8Op and its per-kernel-size launcher (instantiated in strip_gaussq8_k{3,5,7}.cu).
#pragma once

#include "strip_pipeline.cuh"

namespace rcv {

// ---------------------------------------------------------------------------------------
// Op: KS x KS Gaussian (KS = 3, 5, 7) on u8 with arbitrary symmetric Q8 taps (any sigma):
//   out = (sum_i sum_j ky_i kx_j p + 2^15) >> 16      (oracle: orc_sepfilter_u8_q8; == OpenCV)
// Still two samples per register as 16-bit lanes, exactly:
//   vertical   V = sum ky_i p <= 255*256 = 65280 fits a lane;
//   horizontal needs 24 bits, so V is split into bytes Vh = V >> 8, Vl = V & 255 and
//              A = sum kx_j Vh_j, B = sum kx_j Vl_j (each <= 65280) are accumulated separately;
//   (256 A + B + 2^15) >> 16 == (A + (B >> 8) + 128) >> 8, and A + (B >> 8) + 128 <= 65408.
// ---------------------------------------------------------------------------------------
template <int CN, int KS>
struct GaussQ8Op {
  static constexpr int HV = KS / 2;
  static constexpr int P = KS / 2;
  static constexpr int E = CN;
  static constexpr int NOUT = 1;
  static constexpr int WIN = KS == 3 ? 2 : KS == 5 ? 4 : 8;  // window slots (power of two >= KS-1)
  // Rows unrolled in the hot loop = window period (measured: 3x3 69% vs 64% of the roofline at 8).
  // 7x7: an 8-row unroll is a 42 KB loop (> 32 KB I-cache: 44% of stalls were instruction fetch), so
  // its window is shifted physically instead (6 rows x 8 register moves per row) and the loop is 1 row.
  static constexpr bool SHIFT = (KS == 7);
  static constexpr int UNROLL = SHIFT ? 1 : WIN;
  static constexpr int EXT = (HV * CN + 3) / 4;              // neighbour words needed on each side
  static_assert(KS == 3 || KS == 5 || KS == 7, "kernel size");
  static_assert(EXT <= 3, "taps beyond three words");
  uint32_t win[WIN][8];
  uint32_t kx[HV + 1], ky[HV + 1];

  __device__ __forceinline__ void init(const StripParams &p) {
#pragma unroll
    for (int i = 0; i <= HV; ++i) {
      kx[i] = p.taps_x[i];
      ky[i] = p.taps_y[i];
    }
  }
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < WIN; ++j)
#pragma unroll
      for (int h = 0; h < 8; ++h) win[j][h] = 0;
  }

  // symmetric KS-tap sum of packed pairs: t[0..KS-1]
  __device__ __forceinline__ uint32_t sym(const uint32_t (&t)[KS], const uint32_t (&k)[HV + 1]) const {
    uint32_t acc = t[HV] * k[HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) acc += (t[i] + t[KS - 1 - i]) * k[i];
    return acc;
  }

  // neighbour words by shuffle, odd-phase pairs by PRMT, then the KS taps of every output pair
  __device__ __forceinline__ void hpass(const uint32_t (&X)[8], uint32_t (&out)[8]) const {
    constexpr int NWD = 4 + 2 * EXT;  // words -EXT .. 3+EXT at index +EXT
    uint32_t lo[NWD], hi[NWD], loS[NWD - 1], hiS[NWD - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo[k + EXT] = X[2 * k];
      hi[k + EXT] = X[2 * k + 1];
    }
#pragma unroll
    for (int e = 0; e < EXT; ++e) {
      // left lane's words 4-EXT+e -> our word -EXT+e ; right lane's word e -> our word 4+e
      lo[e] = __shfl_up_sync(0xffffffffu, X[2 * (4 - EXT + e)], 1);
      hi[e] = __shfl_up_sync(0xffffffffu, X[2 * (4 - EXT + e) + 1], 1);
      lo[4 + EXT + e] = __shfl_down_sync(0xffffffffu, X[2 * e], 1);
      hi[4 + EXT + e] = __shfl_down_sync(0xffffffffu, X[2 * e + 1], 1);
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

  // SHIFT: win[0] is the oldest row and win[KS-2] the newest; otherwise slot = feed index mod WIN
  __device__ __forceinline__ void push(int slot, const uint32_t (&in)[8]) {
    if (SHIFT) {
#pragma unroll
      for (int i = 0; i < KS - 2; ++i)
#pragma unroll
        for (int h = 0; h < 8; ++h) win[i][h] = win[i + 1][h];
#pragma unroll
      for (int h = 0; h < 8; ++h) win[KS - 2][h] = in[h];
    } else {
#pragma unroll
      for (int h = 0; h < 8; ++h) win[slot][h] = in[h];
    }
  }

  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t in[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[2 * k] = __byte_perm(w[k], 0, 0x4240);
      in[2 * k + 1] = __byte_perm(w[k], 0, 0x4341);
    }
    push(J8 & (WIN - 1), in);
  }

  // J8 = (feed index) & 7, compile time
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    constexpr int J = J8 & (WIN - 1);
    uint32_t in[8];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[2 * k] = __byte_perm(w[k], 0, 0x4240);
      in[2 * k + 1] = __byte_perm(w[k], 0, 0x4341);
    }
    uint32_t Vh[8], Vl[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      uint32_t t[KS];
#pragma unroll
      for (int i = 0; i < KS - 1; ++i) t[i] = SHIFT ? win[i][h] : win[(J8 + WIN * 8 - (KS - 1) + i) & (WIN - 1)][h];  // oldest first
      t[KS - 1] = in[h];
      const uint32_t V = sym(t, ky);  // <= 65280 per lane
      Vl[h] = V & 0x00FF00FFu;
      Vh[h] = (V >> 8) & 0x00FF00FFu;
    }
    push(J, in);
    if (!FAST && !emit) return;
    uint32_t A[8], B[8];
    hpass(Vh, A);
    hpass(Vl, B);
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t H0 = A[2 * k] + ((B[2 * k] >> 8) & 0x00FF00FFu) + 0x00800080u;
      const uint32_t H1 = A[2 * k + 1] + ((B[2 * k + 1] >> 8) & 0x00FF00FFu) + 0x00800080u;
      ow[k] = __byte_perm(H0, H1, 0x7351);
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
  switch (src.v.cn) {
    case 1: return launch_strip<GaussQ8Op<1, KS>>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
    case 2: return launch_strip<GaussQ8Op<2, KS>>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
    case 3: return launch_strip<GaussQ8Op<3, KS>>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
    case 4: return launch_strip<GaussQ8Op<4, KS>>(c, src, &dst, 1, "gauss.band_rows", s, tx, ty);
  }
  return RCV_ERR_UNSUPPORTED;
}



}  // namespace rcv
