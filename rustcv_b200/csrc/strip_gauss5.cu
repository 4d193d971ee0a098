// strip_gauss5.cu -- one op family of the TMA strip pipeline (see strip_pipeline.cuh for the design).
#include "strip_pipeline.cuh"

namespace rcv {

// ---------------------------------------------------------------------------------------
// Op: 5x5 binomial Gaussian on u8, CN interleaved channels.
//   out = (sum_ij k_i k_j p + 128) >> 8, k = {1,4,6,4,1}   (oracle: orc_sepfilter_u8_q8
//   with Q8 taps {16,64,96,64,16}: (sum ky kx p + 32768) >> 16 is the same number)
// ---------------------------------------------------------------------------------------
template <int CN>
struct Gauss5Op {
  static constexpr int HV = 2;   // rows of vertical halo
  static constexpr int P = 2;    // pixels of horizontal halo
  static constexpr int E = CN;   // bytes per pixel
  static constexpr int NOUT = 1;
  static constexpr int UNROLL = 8;       // rows unrolled in the hot loop: the whole chunk (8.05 vs 8.55 us per 4K frame at 4)
  static constexpr int UNROLL_SLOW = 2;  // the per-row-tested loop (ragged strips, a band's last partial chunk) stays small
  static constexpr bool HOIST_WARM = true;
  // Vertical pass in transposed form: instead of the last four rows, the partial sums they have contributed to the
  // next four outputs.  After row x_n:  s4 = x_n,  s3 = x_{n-1} + 4 x_n,  s2 = x_{n-2} + 4 x_{n-1} + 6 x_n,
  // s1 = x_{n-3} + 4 x_{n-2} + 6 x_{n-1} + 4 x_n.  Row x_{n+1} completes V = s1 + x_{n+1} and moves every sum up one
  // slot with one multiply-add each: four operations per register pair like the windowed form, but every state
  // register is rewritten in place, row after row -- no rotating window, so ptxas needs no moves at the loop's
  // back edge (the windowed form carried 8 per row) and the loops may be unrolled by any count.
  uint32_t s1[8], s2[8], s3[8], s4[8];

  __device__ __forceinline__ void init(const StripParams &) {}
  __device__ __forceinline__ void reset() {}  // four warm-up rows overwrite every state register

  // The first 2*HV rows of a band only build the sums: row J8 (0..3) of the band needs J8 operations per pair.
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const uint32_t x = __byte_perm(w[h >> 1], 0, (h & 1) ? 0x4341 : 0x4240);
      if (J8 >= 3) s1[h] = madc<4>(x, s2[h]);
      if (J8 >= 2) s2[h] = madc<6>(x, s3[h]);
      if (J8 >= 1) s3[h] = madc<4>(x, s4[h]);
      s4[h] = x;
    }
  }

  // FAST: interior rows of an aligned, full-width strip -- always emits, lanes store 16 B or nothing.
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    // vertical: V = r0 + 4 r1 + 6 r2 + 4 r3 + r4 (+8 per lane: with horizontal taps summing
    // to 16 that is the final "+128" rounding term).  V <= 4088 per 16-bit lane.
    uint32_t V[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const uint32_t x = __byte_perm(w[h >> 1], 0, (h & 1) ? 0x4341 : 0x4240);  // (b0, b2) / (b1, b3) as 16-bit lanes
      V[h] = add3(s1[h], x, 0x00080008u);
      s1[h] = madc<4>(x, s2[h]);
      s2[h] = madc<6>(x, s3[h]);
      s3[h] = madc<4>(x, s4[h]);
      s4[h] = x;
    }
    if (!FAST && !emit) return;

    // Vertical sums of words -2..5 (index +2) as lo = (byte0, byte2) / hi = (byte1, byte3) pairs and
    // the odd-phase pairs S[i] = (byte 2|3 of word i, byte 0|1 of word i+1).  Own words 0..3; the
    // neighbours' by shuffle: from the left lane its word 3 and its S of words 2-3, from the right
    // lane its word 0 and its S of words 0-1.
    uint32_t lo[8], hi[8], loS[7], hiS[7];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo[k + 2] = V[2 * k];
      hi[k + 2] = V[2 * k + 1];
    }
#pragma unroll
    for (int i = 2; i < 5; ++i) {
      loS[i] = __byte_perm(lo[i], lo[i + 1], 0x5432);
      hiS[i] = __byte_perm(hi[i], hi[i + 1], 0x5432);
    }
    lo[1] = __shfl_up_sync(0xffffffffu, lo[5], 1);
    hi[1] = __shfl_up_sync(0xffffffffu, hi[5], 1);
    loS[0] = __shfl_up_sync(0xffffffffu, loS[4], 1);
    hiS[0] = __shfl_up_sync(0xffffffffu, hiS[4], 1);
    lo[6] = __shfl_down_sync(0xffffffffu, lo[2], 1);
    hi[6] = __shfl_down_sync(0xffffffffu, hi[2], 1);
    loS[6] = __shfl_down_sync(0xffffffffu, loS[2], 1);
    hiS[6] = __shfl_down_sync(0xffffffffu, hiS[2], 1);
    loS[1] = __byte_perm(lo[1], lo[2], 0x5432);
    hiS[1] = __byte_perm(hi[1], hi[2], 0x5432);
    loS[5] = __byte_perm(lo[5], lo[6], 0x5432);
    hiS[5] = __byte_perm(hi[5], hi[6], 0x5432);
    if constexpr (CN == 4) {  // taps at -8 / +8 bytes land on whole words -2 and 5
      lo[0] = __shfl_up_sync(0xffffffffu, lo[4], 1);
      hi[0] = __shfl_up_sync(0xffffffffu, hi[4], 1);
      lo[7] = __shfl_down_sync(0xffffffffu, lo[3], 1);
      hi[7] = __shfl_down_sync(0xffffffffu, hi[3], 1);
    } else {
      lo[0] = hi[0] = lo[7] = hi[7] = 0;  // never selected: CN <= 3 reaches words -2 / 5 only through S[0] / S[6]
    }

    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t H[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        uint32_t t[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int p = 4 * k + e + (j - 2) * CN + 8;  // byte position relative to word -2
          const int wd = p >> 2, ph = p & 3;
          t[j] = ph == 0 ? lo[wd] : ph == 1 ? hi[wd] : ph == 2 ? loS[wd] : hiS[wd];
        }
        H[e] = madc<6>(t[2], madc<4>(add2(t[1], t[3]), add2(t[0], t[4])));  // <= 16 * 4088 = 65408 per lane
      }
      ow[k] = __byte_perm(H[0], H[1], 0x7351);  // high bytes of the four 16-bit lanes, in byte order
    }
    uint8_t *o = outp[0];
    if (FAST) {
      if (nvalid == 16) *(uint4 *)o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else if (nvalid == 16 && vec) {
      *(uint4 *)o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    } else if (nvalid > 0) {  // ragged right edge / unaligned dst only; halo lanes store nothing
#pragma unroll
      for (int b = 0; b < 16; ++b)
        if (b < nvalid) o[b] = (uint8_t)(ow[b >> 2] >> ((b & 3) * 8));
    }
  }
  static_assert(CN >= 1 && CN <= 4, "the farthest taps (2*CN bytes) must stay within two words");
};


// GaussianBlur 5x5 sigma=0 (binomial) fast path.  Returns RCV_ERR_UNSUPPORTED when the
// geometry is not eligible; the caller then uses the generic separable kernel.
int launch_gauss5_strip(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s) {
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_U8) return RCV_ERR_UNSUPPORTED;
  if (src.v.row_bytes() > (size_t)1 << 30) return RCV_ERR_UNSUPPORTED;
  switch (src.v.cn) {
    case 1: return launch_strip<Gauss5Op<1>>(c, src, &dst, 1, "gauss.band_rows", s);
    case 2: return launch_strip<Gauss5Op<2>>(c, src, &dst, 1, "gauss.band_rows", s);
    case 3: return launch_strip<Gauss5Op<3>>(c, src, &dst, 1, "gauss.band_rows", s);
    case 4: return launch_strip<Gauss5Op<4>>(c, src, &dst, 1, "gauss.band_rows", s);
  }
  return RCV_ERR_UNSUPPORTED;
}

}  // namespace rcv
