// strip_f2d_u8.cu -- dense 3x3 filter2D on u8 (1..4 channels) in the TMA strip pipeline.
//   acc = delta; acc = fmaf(k[i][j], (float)p[y+i-1][x+(j-1)*CN], acc) in row-major tap order;
//   out = saturate_u8(rint(acc))           (oracle: orc_filter2d_u8; rint = round half to even)
// Bytes become floats once per source row with the 2^23 trick (PRMT into 0x4B0000bb, minus
// 8388608.0f: exact), the two previous rows stay in registers as floats, and the result is rounded
// by adding 8388608.0f after clamping to [0, 255] (the FADD rounds to nearest even, like rint).
#include "strip_pipeline.cuh"

namespace rcv {

template <int CN>
struct Filter2dU8Op {
  static constexpr int HV = 1;
  static constexpr int P = 1;
  static constexpr int E = CN;
  static constexpr int NOUT = 1;
  static constexpr int UNROLL = 2;
  static constexpr int XW = 16 + 2 * CN;  // element columns -CN .. 15+CN
  float win[2][XW];
  float k[9];
  float delta;

  __device__ __forceinline__ void init(const StripParams &p) {
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = p.ftaps[i];
    delta = p.ftaps[9];
  }
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < XW; ++h) win[j][h] = 0.0f;
  }
  static __device__ __forceinline__ float byte_to_float(uint32_t w, int b) {
    // [byte b of w, 0, 0, 0x4B] = 8388608 + byte as a float
    const uint32_t sel = b == 0 ? 0x7650u : b == 1 ? 0x7651u : b == 2 ? 0x7652u : 0x7653u;
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)), 8388608.0f);
  }
  __device__ __forceinline__ void widen(const uint4 &q, float (&x)[XW]) const {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    const uint32_t wl = __shfl_up_sync(0xffffffffu, w[3], 1);    // left lane's bytes 12..15
    const uint32_t wr = __shfl_down_sync(0xffffffffu, w[0], 1);  // right lane's bytes 0..3
#pragma unroll
    for (int e = 0; e < CN; ++e) {
      x[e] = byte_to_float(wl, 4 - CN + e);
      x[CN + 16 + e] = byte_to_float(wr, e);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) x[CN + e] = byte_to_float(w[e >> 2], e & 3);
  }
  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {
    float x[XW];
    widen(q, x);
#pragma unroll
    for (int c = 0; c < XW; ++c) win[J8 & 1][c] = x[c];
  }
  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    constexpr int J = J8 & 1;
    float x[XW];
    widen(q, x);
    uint32_t ow[4] = {0, 0, 0, 0};
    if (FAST || emit) {
      uint32_t r[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float acc = delta;
#pragma unroll
        for (int j = 0; j < 3; ++j) acc = fmaf(k[j], win[J][e + j * CN], acc);  // oldest row (y-1)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc = fmaf(k[3 + j], win[J ^ 1][e + j * CN], acc);
#pragma unroll
        for (int j = 0; j < 3; ++j) acc = fmaf(k[6 + j], x[e + j * CN], acc);
        acc = fminf(fmaxf(acc, 0.0f), 255.0f);
        r[e] = __float_as_uint(__fadd_rn(acc, 8388608.0f));  // low byte = rint(acc)
      }
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const uint32_t lo = __byte_perm(r[4 * w4], r[4 * w4 + 1], 0x0040);
        const uint32_t hi = __byte_perm(r[4 * w4 + 2], r[4 * w4 + 3], 0x0040);
        ow[w4] = __byte_perm(lo, hi, 0x5410);
      }
    }
#pragma unroll
    for (int c = 0; c < XW; ++c) win[J][c] = x[c];
    if (!FAST && !emit) return;
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

int launch_filter2d_u8_strip(Ctx *c, const DBatch &src, const DBatch &dst, const float *k, int kw, int kh, float delta,
                             cudaStream_t s) {
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_U8 || kw != 3 || kh != 3) return RCV_ERR_UNSUPPORTED;
  if (src.v.row_bytes() > (size_t)1 << 30) return RCV_ERR_UNSUPPORTED;
  float taps[10];
  for (int i = 0; i < 9; ++i) taps[i] = k[i];
  taps[9] = delta;
  switch (src.v.cn) {
    case 1: return launch_strip<Filter2dU8Op<1>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, 10);
    case 2: return launch_strip<Filter2dU8Op<2>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, 10);
    case 3: return launch_strip<Filter2dU8Op<3>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, 10);
    case 4: return launch_strip<Filter2dU8Op<4>>(c, src, &dst, 1, "f2d.band_rows", s, nullptr, nullptr, taps, 10);
  }
  return RCV_ERR_UNSUPPORTED;
}

}  // namespace rcv
