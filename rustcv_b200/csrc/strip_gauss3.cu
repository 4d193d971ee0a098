// strip_gauss3.cu -- one op family of the TMA strip pipeline (see strip_pipeline.cuh for the design).
#include "strip_pipeline.cuh"

namespace rcv {

// ---------------------------------------------------------------------------------------
// Op: 3x3 binomial Gaussian on u8 (GaussianBlur ksize 3, sigma <= 0: taps {1,2,1}/4), CN interleaved channels.
//   out = (sum_ij b_i b_j p + 8) >> 4                (oracle: orc_sepfilter_u8_q8 with Q8 taps {64,128,64}:
//                                                     (4096 S + 2^15) >> 16 is the same number)
// The sibling of Gauss5Op: two samples per register as 16-bit lanes.  Vertical V = r0 + 2 r1 + r2 <= 1020;
// horizontal H = 16 (V[-1] + 2 V[0] + V[+1]) + 128 <= 65408, scaled by 16 so that the result byte is the HIGH
// byte of each lane ((16 S + 128) >> 8 == (S + 8) >> 4) and one PRMT packs four of them.  5 integer ops per
// register pair against the any-sigma op's ~14: the 3x3 default blur becomes memory bound like the 5x5 one.
// ---------------------------------------------------------------------------------------
template <int CN>
struct Gauss3Op {
  static constexpr int HV = 1;
  static constexpr int P = 1;
  static constexpr int E = CN;
  static constexpr int NOUT = 1;
  static constexpr int UNROLL = 8;  // window period 2; whole chunks unrolled like Gauss5Op
  static constexpr bool HOIST_WARM = true;  // like Gauss5Op: a straight-line copy of the band's first chunk
  uint32_t win[2][8];               // last 2 rows, unpacked: [2w] = bytes 0,2 of word w; [2w+1] = bytes 1,3

  __device__ __forceinline__ void init(const StripParams &) {}
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 8; ++h) win[j][h] = 0;
  }

  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      win[J8 & 1][2 * k] = __byte_perm(w[k], 0, 0x4240);
      win[J8 & 1][2 * k + 1] = __byte_perm(w[k], 0, 0x4341);
    }
  }

  // J = (feed index) & 1, compile time: win[J] holds the older row.
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    constexpr int J = J8 & 1;
    uint32_t in[8];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[2 * k] = __byte_perm(w[k], 0, 0x4240);      // (b0, b2) as 16-bit lanes
      in[2 * k + 1] = __byte_perm(w[k], 0, 0x4341);  // (b1, b3)
    }
    uint32_t V[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      V[h] = madc<2>(win[J ^ 1][h], add2(win[J][h], in[h]));  // <= 1020 per lane
      win[J][h] = in[h];
    }
    if (!FAST && !emit) return;

    // words -1..4 at index +2 (the layout of Gauss5Op): own words 0..3, the neighbours' by shuffle; odd-phase
    // pairs S[i] = (byte 2|3 of word i, byte 0|1 of word i+1)
    uint32_t lo[8], hi[8], loS[7], hiS[7];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo[k + 2] = V[2 * k];
      hi[k + 2] = V[2 * k + 1];
    }
    lo[1] = __shfl_up_sync(0xffffffffu, lo[5], 1);
    hi[1] = __shfl_up_sync(0xffffffffu, hi[5], 1);
    lo[6] = __shfl_down_sync(0xffffffffu, lo[2], 1);
    hi[6] = __shfl_down_sync(0xffffffffu, hi[2], 1);
    lo[0] = hi[0] = lo[7] = hi[7] = 0;  // never selected
#pragma unroll
    for (int i = 1; i < 6; ++i) {
      loS[i] = __byte_perm(lo[i], lo[i + 1], 0x5432);
      hiS[i] = __byte_perm(hi[i], hi[i + 1], 0x5432);
    }
    loS[0] = hiS[0] = loS[6] = hiS[6] = 0;

    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t H[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        uint32_t t[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int p = 4 * k + e + (j - 1) * CN + 8;  // byte position relative to word -2
          const int wd = p >> 2, ph = p & 3;
          t[j] = ph == 0 ? lo[wd] : ph == 1 ? hi[wd] : ph == 2 ? loS[wd] : hiS[wd];
        }
        H[e] = madc<16>(madc<2>(t[1], add2(t[0], t[2])), 0x00800080u);  // 16 S + 128 <= 65408 per lane
      }
      ow[k] = __byte_perm(H[0], H[1], 0x7351);  // high bytes of the four 16-bit lanes, in byte order
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
  static_assert(CN >= 1 && CN <= 4, "the taps (CN bytes away) must stay within the neighbouring word");
};

// GaussianBlur 3x3 sigma<=0 (binomial) fast path.  RCV_ERR_UNSUPPORTED -> the caller uses the any-sigma op.
int launch_gauss3_strip(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s) {
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_U8) return RCV_ERR_UNSUPPORTED;
  if (src.v.row_bytes() > (size_t)1 << 30) return RCV_ERR_UNSUPPORTED;
  switch (src.v.cn) {
    case 1: return launch_strip<Gauss3Op<1>>(c, src, &dst, 1, "gauss.band_rows", s);
    case 2: return launch_strip<Gauss3Op<2>>(c, src, &dst, 1, "gauss.band_rows", s);
    case 3: return launch_strip<Gauss3Op<3>>(c, src, &dst, 1, "gauss.band_rows", s);
    case 4: return launch_strip<Gauss3Op<4>>(c, src, &dst, 1, "gauss.band_rows", s);
  }
  return RCV_ERR_UNSUPPORTED;
}

}  // namespace rcv
