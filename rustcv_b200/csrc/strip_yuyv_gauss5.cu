// strip_yuyv_gauss5.cu -- fused decode -> process chain on the TMA strip pipeline:
//
//     dst = GaussianBlur5x5( YUYV2BGR(src) )        (sigma = 0: binomial taps, BORDER_REFLECT_101)
//
// in ONE kernel: 2 B/px read (the raw camera frame, rustcv/src/videoio/mod.rs:201-205), 3 B/px written = 5 B/px;
// the conversion kernel followed by the blur kernel moves 2+3 and 3+3 = 11 B/px.  Both stages are the
// arithmetic of the stand-alone ops, so the result is bit-identical to the chain:
//   YUYV2BGR  the reference's BT.601 integer formula (videoio/mod.rs:344-371), cvt_math.cuh
//   blur      out = (sum_ij k_i k_j p + 128) >> 8, k = {1,4,6,4,1}: exact integer, single rounding (Gauss5Op)
//
// Layout.  A lane owns 16 input bytes = 4 macro-pixels = 8 pixels and writes 24 output bytes (OMUL/ODIV = 3/2).
// The conversion hands out the two pixels of a macro-pixel per channel, which is exactly the packed form the
// blur wants: one register = (pixel 2k, pixel 2k+1) of ONE channel as 16-bit lanes -- planar in registers, so
// the horizontal taps are whole-register neighbours (P) and their odd-phase pairs (Q = PRMT 0x5432) instead
// of the interleaved stride-3 byte gathers of Gauss5Op<3>.  The window keeps the converted rows (4 x 12
// registers); vertical sums V <= 4088 and horizontal sums H <= 65408 fit the 16-bit lanes exactly.
//
// Borders.  Rows reflect on the raw bytes in shared memory like every op (the conversion is row-local).  Columns
// cannot: BGR pixel -2 is pixel 2 and -1 is pixel 1, which come from two different macro-pixels (two chroma
// pairs), so no YUYV macro-pixel at position -1 can stand for both.  The op therefore reflects its VERTICAL
// SUMS in registers (Op::EDGES): with the last valid pixel of the row at local index 2m+1 of a lane,
// P[m+1] = swap(Q[m-1]) and Q[m] = swap(P[m]); at the left edge P[-1] = swap(Q[0]) and Q[-1] = swap(P[0]).
#include "cvt_math.cuh"
#include "strip_pipeline.cuh"

namespace rcv {

struct YuyvGauss5Op {
  static constexpr int HV = 2;
  static constexpr int P = 0;  // no shared-memory column patches: see EDGES
  static constexpr int E = 4;
  static constexpr bool EDGES = true;
  static constexpr int OMUL = 3, ODIV = 2;
  static constexpr int BAND_ROWS = 60;  // measured best of 28..124 (heavier rows: fewer re-converted warm-up rows)
  static constexpr int NOUT = 1;
  // Vertical pass in transposed form (see Gauss5Op): partial sums instead of a rotating window of rows, so the loops
  // need no particular unroll count.  Two rows per iteration keep the steady loop (~240 instructions per row), the
  // per-row-tested loop and the four straight-line warm-up rows together inside the 32 KB instruction cache -- with the
  // 4-row windowed form only ONE (predicated) copy of the row loop fitted (r1 ncu capture: stall_no_inst 20 %).
  static constexpr int UNROLL = 2;
  static constexpr bool HOIST_WARM = true;
  uint32_t s1[12], s2[12], s3[12], s4[12];  // [3k + c] = channel c (B, G, R) of macro-pixel k, see Gauss5Op
  int left_lane, edge_lane, edge_m;  // lane to patch at the left edge / right edge (-1: none), m of the right edge

  __device__ __forceinline__ void init(const StripParams &) {}
  __device__ __forceinline__ void reset() {}  // four warm-up rows overwrite every partial sum

  // xr = tile byte offset of the first byte past the row (16 = start of lane 1)
  __device__ __forceinline__ void edges(bool left, bool right, int xr, int) {
    left_lane = left ? 1 : -1;
    edge_lane = -1;
    edge_m = 0;
    if (right) {
      const int last = xr - 4;  // tile byte offset of the row's last macro-pixel
      edge_lane = last >> 4;
      edge_m = (last & 15) >> 2;
    }
  }

  static __device__ __forceinline__ uint32_t swap16(uint32_t v) { return __byte_perm(v, 0, 0x1032); }

  // 4 macro-pixels -> 12 packed registers: (value of pixel 2k, value of pixel 2k+1) per channel
  static __device__ __forceinline__ void convert(const uint4 &q, uint32_t (&o)[12]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) yuv_word_pairs<false>(w[k], o[3 * k + 0], o[3 * k + 1], o[3 * k + 2]);
  }

  template <int J8>
  __device__ __forceinline__ void warm(const uint4 &q) {  // J8 = row of the band (0..3): J8 operations per pair
    uint32_t in[12];
    convert(q, in);
#pragma unroll
    for (int h = 0; h < 12; ++h) {
      if (J8 >= 3) s1[h] = madc<4>(in[h], s2[h]);
      if (J8 >= 2) s2[h] = madc<6>(in[h], s3[h]);
      if (J8 >= 1) s3[h] = madc<4>(in[h], s4[h]);
      s4[h] = in[h];
    }
  }

  template <int J8, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    uint32_t in[12];
    convert(q, in);
    // vertical: V = r0 + 4 r1 + 6 r2 + 4 r3 + r4 (+8 per lane = the final +128 after the 16-weight row pass)
    uint32_t V[12];
#pragma unroll
    for (int h = 0; h < 12; ++h) {
      V[h] = add3(s1[h], in[h], 0x00080008u);
      s1[h] = madc<4>(in[h], s2[h]);
      s2[h] = madc<6>(in[h], s3[h]);
      s3[h] = madc<4>(in[h], s4[h]);
      s4[h] = in[h];
    }
    if (!FAST && !emit) return;  // warm-up rows of a band only build the sums
    const int lane = threadIdx.x & 31;
    uint32_t Hc[3][4];
    uint32_t Pp[3][6], Qq[3][5];  // per channel: P[-1..4] at index +1, Q[-1..3] at index +1
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 4; ++k) Pp[c][k + 1] = V[3 * k + c];
      Pp[c][0] = __shfl_up_sync(0xffffffffu, Pp[c][4], 1);
      Pp[c][5] = __shfl_down_sync(0xffffffffu, Pp[c][1], 1);
#pragma unroll
      for (int k = 0; k < 5; ++k) Qq[c][k] = __byte_perm(Pp[c][k], Pp[c][k + 1], 0x5432);
    }
    if (left_lane >= 0 || edge_lane >= 0) {  // warp-uniform: edge strips only (one branch per row).  Selects, not indexed stores.
      const bool el = lane == left_lane, er = lane == edge_lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t lp = swap16(Qq[c][1]), lq = swap16(Pp[c][1]);
        Pp[c][0] = el ? lp : Pp[c][0];
        Qq[c][0] = el ? lq : Qq[c][0];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const bool hit = er && edge_m == m;
          const uint32_t np = swap16(Qq[c][m]);      // P[m+1] = swap(Q[m-1])
          const uint32_t nq = swap16(Pp[c][m + 1]);  // Q[m]   = swap(P[m])
          Pp[c][m + 2] = hit ? np : Pp[c][m + 2];
          Qq[c][m + 1] = hit ? nq : Qq[c][m + 1];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        Hc[c][k] = madc<6>(Pp[c][k + 1], madc<4>(add2(Qq[c][k], Qq[c][k + 1]), add2(Pp[c][k], Pp[c][k + 2])));  // <= 65408 per lane
    // pack: the result bytes are the high bytes of the 16-bit lanes; two macro-pixels -> 12 bytes = 3 words
    uint32_t ow[6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t X = __byte_perm(Hc[1][2 * h], Hc[2][2 * h], 0x7351);          // G0 R0 G1 R1
      const uint32_t Y = __byte_perm(Hc[0][2 * h + 1], Hc[1][2 * h + 1], 0x7351);  // B2 G2 B3 G3
      ow[3 * h + 0] = __byte_perm(Hc[0][2 * h], X, 0x3541);                        // B0 G0 R0 B1
      ow[3 * h + 1] = __byte_perm(X, Y, 0x5432);                                   // G1 R1 B2 G2
      ow[3 * h + 2] = __byte_perm(Hc[2][2 * h + 1], Y, 0x3761);                    // R2 B3 G3 R3
    }
    uint8_t *o = outp[0];
    if (FAST) {
      if (nvalid == 24) {
        *(uint2 *)o = make_uint2(ow[0], ow[1]);
        *(uint2 *)(o + 8) = make_uint2(ow[2], ow[3]);
        *(uint2 *)(o + 16) = make_uint2(ow[4], ow[5]);
      }
    } else if (nvalid == 24 && vec) {
      *(uint2 *)o = make_uint2(ow[0], ow[1]);
      *(uint2 *)(o + 8) = make_uint2(ow[2], ow[3]);
      *(uint2 *)(o + 16) = make_uint2(ow[4], ow[5]);
    } else if (nvalid > 0) {
#pragma unroll
      for (int b = 0; b < 24; ++b)
        if (b < nvalid) o[b] = (uint8_t)(ow[b >> 2] >> ((b & 3) * 8));
    }
  }
};

// src: YUYV as a 2-channel u8 image (even cols >= 8); dst: 3-channel u8 of the same rows x cols.
int launch_yuyv_gauss5_strip(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s) {
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_U8 || src.v.cn != 2 || (src.v.cols & 1)) return RCV_ERR_UNSUPPORTED;
  if (!dst.v.data || dst.v.depth != RCV_U8 || dst.v.cn != 3) return RCV_ERR_UNSUPPORTED;
  return launch_strip<YuyvGauss5Op>(c, src, &dst, 1, "yuyvgauss.band_rows", s);
}

}  // namespace rcv
