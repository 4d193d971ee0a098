// geom.cu -- bilinear resize and warpAffine.
//
// Absent from the reference (the only "resize" is a vendor call,
// rustcv-camera/src/backend/macos/bridge.m:140); semantics are the oracle's
// (oracle/rcv_oracle.c: orc_resize_bilinear_{u8,f32}, orc_warp_affine_{f32,u8}):
//   resize u8   cv::resize INTER_LINEAR fixed point (11-bit weights), bit-exact with OpenCV
//   resize f32  h = p0*(1-fx) + p1*fx ; out = h0*(1-fy) + h1*fy, each op rounded once
//   warpAffine  inverse map, f64 row term, fmaf column term, fmaf lerp chain, constant border
// The general resize is a per-pixel gather kernel (a shared-memory staged tile variant was built and measured
// SLOWER on single 4K frames -- 13.7 vs 9.0 us for 4K->720p, 21.3 vs 15.0 us for 4K->1080p: at these sizes the
// kernel is launch/ramp bound and the two-phase tile adds a barrier -- and was removed); the exact 4x BGR
// downscale (BASELINE.json config 4) has a 128-bit-load specialisation that runs at the DRAM sector floor.
#include "rcv_internal.cuh"
#include "tma_ptx.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace rcv {

// ---------------------------------------------------------------------------------------
// resize: per-column / per-row tables built on the host exactly as the oracle does
// ---------------------------------------------------------------------------------------
struct ResizeCol {
  int x0, x1;      // source columns (clamped)
  int a0, a1;      // u8: 11-bit fixed-point weights
  float f0, f1;    // f32 weights (1-fx, fx)
};
struct ResizeRow {
  int y0, y1;
  int b0, b1;
  float f0, f1;
};

static void resize_tables(int srows, int scols, int drows, int dcols, std::vector<ResizeCol> &cols,
                          std::vector<ResizeRow> &rows) {
  const double scale_x = (double)scols / dcols, scale_y = (double)srows / drows;
  cols.resize(dcols);
  rows.resize(drows);
  for (int dx = 0; dx < dcols; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) {
      fx = 0.0f;
      sx = 0;
    }
    if (sx >= scols - 1) {
      fx = 0.0f;
      sx = scols - 1;
    }
    ResizeCol c;
    c.x0 = sx;
    c.x1 = sx + 1 < scols ? sx + 1 : sx;
    c.a0 = (short)lrintf((1.0f - fx) * 2048.0f);
    c.a1 = (short)lrintf(fx * 2048.0f);
    c.f0 = 1.0f - fx;
    c.f1 = fx;
    cols[dx] = c;
  }
  for (int dy = 0; dy < drows; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= (float)sy;
    ResizeRow r;
    r.y0 = sy < 0 ? 0 : (sy >= srows ? srows - 1 : sy);
    r.y1 = sy + 1 < 0 ? 0 : (sy + 1 >= srows ? srows - 1 : sy + 1);
    r.b0 = (short)lrintf((1.0f - fy) * 2048.0f);
    r.b1 = (short)lrintf(fy * 2048.0f);
    r.f0 = 1.0f - fy;
    r.f1 = fy;
    rows[dy] = r;
  }
}

struct ResizeArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *dst;
  size_t dstep, dfs;
  int srows, scols, drows, dcols, cn;
  const ResizeCol *cols;
  const ResizeRow *rows;
  const int2 *cols8;  // compact u8 column table: .x = x0, .y = a0 | a1 << 16 (k_resize_u8w)
};

template <typename T>
__global__ void __launch_bounds__(256) k_resize_generic(const ResizeArgs a) {
  int dx = blockIdx.x * blockDim.x + threadIdx.x;
  int dy = blockIdx.y;
  if (dx >= a.dcols) return;
  const ResizeCol c = a.cols[dx];
  const ResizeRow r = a.rows[dy];
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const T *s0 = (const T *)(src + (size_t)r.y0 * a.sstep);
  const T *s1 = (const T *)(src + (size_t)r.y1 * a.sstep);
  T *d = (T *)(a.dst + (size_t)blockIdx.z * a.dfs + (size_t)dy * a.dstep);
  for (int ch = 0; ch < a.cn; ++ch) {
    if (sizeof(T) == 1) {
      int S0 = (int)s0[c.x0 * a.cn + ch] * c.a0 + (int)s0[c.x1 * a.cn + ch] * c.a1;
      int S1 = (int)s1[c.x0 * a.cn + ch] * c.a0 + (int)s1[c.x1 * a.cn + ch] * c.a1;
      int v = (((r.b0 * (S0 >> 4)) >> 16) + ((r.b1 * (S1 >> 4)) >> 16) + 2) >> 2;
      d[dx * a.cn + ch] = (T)min(max(v, 0), 255);
    } else {
      float t0 = __fmul_rn((float)s0[c.x0 * a.cn + ch], c.f0);
      float t1 = __fmul_rn((float)s0[c.x1 * a.cn + ch], c.f1);
      float h0 = __fadd_rn(t0, t1);
      float t2 = __fmul_rn((float)s1[c.x0 * a.cn + ch], c.f0);
      float t3 = __fmul_rn((float)s1[c.x1 * a.cn + ch], c.f1);
      float h1 = __fadd_rn(t2, t3);
      d[dx * a.cn + ch] = (T)__fadd_rn(__fmul_rn(h0, r.f0), __fmul_rn(h1, r.f1));
    }
  }
}

// Exact 4x downscale of 3-channel u8 (config 4: 7680x4320 -> 1920x1080).  With scale 4 the
// oracle's weights are all 1024 and its fixed-point chain collapses to
//   out = (p[4y+1][4x+1] + p[4y+1][4x+2] + p[4y+2][4x+1] + p[4y+2][4x+2] + 2) >> 2
// (checked bit-exact against the general model in tests).
// A thread makes 4 dst pixels: 2 rows x 48 B in (3 x LDG.128 each), 12 B out.
__device__ __forceinline__ uint32_t byte_at(const uint32_t *w, int k) { return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu; }

__global__ void __launch_bounds__(128) k_resize4x_u8c3(const ResizeArgs a) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 dst pixels
  int dy = blockIdx.y;
  int groups = a.dcols >> 2;
  if (g >= groups) return;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const uint4 *r1 = (const uint4 *)(src + (size_t)(4 * dy + 1) * a.sstep + (size_t)g * 48);
  const uint4 *r2 = (const uint4 *)(src + (size_t)(4 * dy + 2) * a.sstep + (size_t)g * 48);
  uint32_t w1[12], w2[12];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    uint4 q = __ldg(r1 + k);
    w1[4 * k] = q.x;
    w1[4 * k + 1] = q.y;
    w1[4 * k + 2] = q.z;
    w1[4 * k + 3] = q.w;
    q = __ldg(r2 + k);
    w2[4 * k] = q.x;
    w2[4 * k + 1] = q.y;
    w2[4 * k + 2] = q.z;
    w2[4 * k + 3] = q.w;
  }
  uint32_t out[3] = {0, 0, 0};
#pragma unroll
  for (int ob = 0; ob < 12; ++ob) {
    const int j = ob / 3, ch = ob % 3;
    const int i0 = 12 * j + 3 + ch, i1 = i0 + 3;
    uint32_t v = (byte_at(w1, i0) + byte_at(w1, i1) + byte_at(w2, i0) + byte_at(w2, i1) + 2u) >> 2;
    out[ob >> 2] |= v << ((ob & 3) * 8);
  }
  uint32_t *d = (uint32_t *)(a.dst + (size_t)blockIdx.z * a.dfs + (size_t)dy * a.dstep + (size_t)g * 12);
  d[0] = out[0];
  d[1] = out[1];
  d[2] = out[2];
}

// General resize of u8 C3 / C4, word loads.  k_resize_generic is bound by its load/store unit traffic (12 byte
// loads and 3 byte stores per BGR pixel, each warp-wide load two wavefronts wide).  The two taps of a row are
// adjacent pixels, i.e. 2*CN contiguous bytes: here they are fetched as the aligned 32-bit words that cover them
// and moved into place with funnel shifts (6 word loads per BGR pixel instead of 12 byte loads), and a thread
// makes 4 destination pixels so that its output leaves as words.  Same fixed-point arithmetic.
// One destination pixel of k_resize_u8w.  SAFE: the words wi .. wi+2 of both rows exist (every pixel but the last
// few of a row), so the three loads of a row share one 64-bit address with immediate offsets; otherwise the word index
// is clamped to the row's last word per load (a 64-bit address computation each).
template <int CN, bool SAFE>
__device__ __forceinline__ void resize_px_u8w(const uint32_t *s0, const uint32_t *s1, const int2 e, int last_word, uint32_t b0s,
                                              uint32_t b1s, uint32_t *ob) {
  const int x0 = e.x;
  const uint32_t a0 = (uint32_t)e.y & 0xFFFFu, a1 = (uint32_t)e.y >> 16;
  const int o = x0 * CN, wi = o >> 2, sh = (o & 3) * 8;
  uint32_t p0[2][CN], p1[2][CN];  // [row][channel]: tap (x0) and tap (x1)
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const uint32_t *row = (rr ? s1 : s0) + wi;
    // the two taps are bytes o .. o + 2*CN - 1: one to three aligned words, shifted into place
    const uint32_t w0 = __ldg(row);
    uint32_t lo, hi = 0;
    if (CN == 4) {
      lo = w0;  // o is a multiple of 4: the two pixels are the two words
      hi = SAFE ? __ldg(row + 1) : __ldg(row + min(1, last_word - wi));
    } else if (CN == 3) {
      const uint32_t w1 = SAFE ? __ldg(row + 1) : __ldg(row + min(1, last_word - wi));
      uint32_t w2 = 0;  // a third word only when the first tap starts at byte 3 of its word
      if (sh == 24) w2 = SAFE ? __ldg(row + 2) : __ldg(row + min(2, last_word - wi));
      lo = __funnelshift_r(w0, w1, sh);  // bytes o .. o+3
      hi = __funnelshift_r(w1, w2, sh);  // bytes o+4 .. o+7
    } else {
      uint32_t w1 = 0;  // gray / 2 channels: the second word only when the taps straddle a word boundary
      if (sh + 16 * CN > 32) w1 = SAFE ? __ldg(row + 1) : __ldg(row + min(1, last_word - wi));
      lo = __funnelshift_r(w0, w1, sh);
    }
    // one PRMT per tap byte (shift + mask were two ALU-pipe operations each, and that pipe is this kernel's limit:
    // profiles/r2_resize_general_ncu_before.txt, ALU 85 % busy).  A column clamped at the right edge has
    // a0 = 2048, a1 = 0 in the compact table, so its second tap (whatever lies there) contributes nothing.
#pragma unroll
    for (int ch = 0; ch < CN; ++ch) {
      p0[rr][ch] = __byte_perm(lo, 0, 0x4440 | ch);
      const int k = CN + ch;  // byte index of the second tap
      p1[rr][ch] = k < 4 ? __byte_perm(lo, 0, 0x4440 | k) : __byte_perm(hi, 0, 0x4440 | (k - 4));
    }
  }
#pragma unroll
  for (int ch = 0; ch < CN; ++ch) {
    const uint32_t S0 = p0[0][ch] * a0 + p1[0][ch] * a1;
    const uint32_t S1 = p0[1][ch] * a0 + p1[1][ch] * a1;
    // (b * (S >> 4)) >> 16 as the high word of (b << 16) * (S >> 4): one multiply, no shift
    const uint32_t v = (__umulhi(S0 >> 4, b0s) + __umulhi(S1 >> 4, b1s) + 2u) >> 2;
    ob[ch] = min(v, 255u);  // v <= 255 whenever the weights sum to 2048; kept for the oracle's saturate
  }
}

template <int CN>
__global__ void __launch_bounds__(128) k_resize_u8w(const ResizeArgs a) {
  static_assert(CN >= 1 && CN <= 4, "1..4 interleaved u8 channels");
  // block = 32 groups x 4 rows: a warp is 32 consecutive groups of one row, and a row of G groups wastes at most
  // 31 threads (128 groups x 1 row left 1280- and 1600-column rows with 17 % / 22 % of the threads idle)
  const int t = blockIdx.x * 32 + (threadIdx.x & 31);  // group of 4 destination pixels
  const int dy = blockIdx.y * 4 + (threadIdx.x >> 5);
  const int dx0 = 4 * t;
  if (dx0 >= a.dcols || dy >= a.drows) return;
  const ResizeRow r = a.rows[dy];
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const uint32_t *s0 = (const uint32_t *)(src + (size_t)r.y0 * a.sstep);
  const uint32_t *s1 = (const uint32_t *)(src + (size_t)r.y1 * a.sstep);
  const int last_word = (a.scols * CN - 1) >> 2;  // the last word holding row bytes
  const uint32_t b0s = (uint32_t)r.b0 << 16, b1s = (uint32_t)r.b1 << 16;
  uint32_t ob[4 * CN];                            // output bytes of the 4 pixels
  // the group's four 8-byte table entries (x0, a0 | a1 << 16) as two 128-bit loads: the table is padded with copies of
  // its last entry, so a ragged last group recomputes its last pixel (not stored).  (The 24-byte ResizeCol cost a warp
  // 12 L1 wavefronts per pixel column.)
  const int4 t0 = __ldg((const int4 *)(a.cols8 + dx0)), t1 = __ldg((const int4 *)(a.cols8 + dx0) + 1);
  const int2 e[4] = {make_int2(t0.x, t0.y), make_int2(t0.z, t0.w), make_int2(t1.x, t1.y), make_int2(t1.z, t1.w)};
  if (((e[3].x * CN) >> 2) + 2 <= last_word) {  // x0 grows with dx: the last pixel's words exist, so all do
#pragma unroll
    for (int q = 0; q < 4; ++q) resize_px_u8w<CN, true>(s0, s1, e[q], last_word, b0s, b1s, ob + q * CN);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) resize_px_u8w<CN, false>(s0, s1, e[q], last_word, b0s, b1s, ob + q * CN);
  }
  uint8_t *d = a.dst + (size_t)blockIdx.z * a.dfs + (size_t)dy * a.dstep + (size_t)dx0 * CN;
  if (dx0 + 4 <= a.dcols) {
    uint32_t *dw = (uint32_t *)d;  // 4*CN bytes, 4-byte aligned (dst base and step are)
#pragma unroll
    for (int k = 0; k < CN; ++k)
      dw[k] = ob[4 * k] | (ob[4 * k + 1] << 8) | (ob[4 * k + 2] << 16) | (ob[4 * k + 3] << 24);
  } else {
    const int n = (a.dcols - dx0) * CN;
#pragma unroll
    for (int k = 0; k < 4 * CN; ++k)
      if (k < n) d[k] = (uint8_t)ob[k];
  }
}

// Exact 2x downscale of u8 (4K -> 1080p, 1080p -> 540p: the commonest resize).  With scale 2 every weight of
// the oracle's fixed-point model is 1024 and its chain collapses, exactly, to the rounded 2x2 box mean
//   out = (p[2y][2x] + p[2y][2x+1] + p[2y+1][2x] + p[2y+1][2x+1] + 2) >> 2
// ((1024 * ((1024 (a + b)) >> 4)) >> 16 == a + b; tested equal to the general kernel and to OpenCV).  Every source
// byte is needed, so the kernel is a pure stream: a thread makes G dst pixels from 2 rows x 2*G*CN bytes
// (128-bit loads), G*CN bytes out.
template <int CN, int G>
__global__ void __launch_bounds__(128) k_resize2x_u8(const ResizeArgs a) {
  constexpr int IN_B = 2 * G * CN, OUT_B = G * CN;
  static_assert(IN_B % 16 == 0 && OUT_B % 4 == 0, "vector widths");
  const int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of G dst pixels
  const int dy = blockIdx.y;
  if (g >= a.dcols / G) return;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const uint4 *r0 = (const uint4 *)(src + (size_t)(2 * dy) * a.sstep + (size_t)g * IN_B);
  const uint4 *r1 = (const uint4 *)(src + (size_t)(2 * dy + 1) * a.sstep + (size_t)g * IN_B);
  uint32_t w0[IN_B / 4], w1[IN_B / 4];
#pragma unroll
  for (int k = 0; k < IN_B / 16; ++k) {
    uint4 q = __ldg(r0 + k);
    w0[4 * k] = q.x;
    w0[4 * k + 1] = q.y;
    w0[4 * k + 2] = q.z;
    w0[4 * k + 3] = q.w;
    q = __ldg(r1 + k);
    w1[4 * k] = q.x;
    w1[4 * k + 1] = q.y;
    w1[4 * k + 2] = q.z;
    w1[4 * k + 3] = q.w;
  }
  uint32_t out[OUT_B / 4];
#pragma unroll
  for (int k = 0; k < OUT_B / 4; ++k) out[k] = 0;
#pragma unroll
  for (int ob = 0; ob < OUT_B; ++ob) {
    const int j = ob / CN, ch = ob % CN;
    const int i0 = 2 * j * CN + ch, i1 = i0 + CN;
    const uint32_t v = (byte_at(w0, i0) + byte_at(w0, i1) + byte_at(w1, i0) + byte_at(w1, i1) + 2u) >> 2;
    out[ob >> 2] |= v << ((ob & 3) * 8);
  }
  uint32_t *d = (uint32_t *)(a.dst + (size_t)blockIdx.z * a.dfs + (size_t)dy * a.dstep + (size_t)g * OUT_B);
#pragma unroll
  for (int k = 0; k < OUT_B / 4; ++k) d[k] = out[k];
}

template <int CN, int G>
static int launch_resize2x(const ResizeArgs &a, int n, cudaStream_t s) {
  dim3 grid(ceil_div(a.dcols / G, 128), a.drows, n);
  k_resize2x_u8<CN, G><<<grid, 128, 0, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

// Exact 2x upscale of u8 (1080p -> 4K).  With scale 1/2 the oracle's weights are 512 / 1536 (and 2048 / 0 at
// the clamped edge columns), and its fixed-point chain collapses, exactly, to
//   H(row, dst col)  = c[far] + 3 c[near]          far = the source pixel on the other side, clamped to the row
//   out(dst row, .)  = ((H[far row] >> 2) + ((3 H[near row]) >> 2) + 2) >> 2      rows clamped likewise
// (checked against the oracle and OpenCV for all channel counts, tests + the derivation in DESIGN.md).  A thread
// takes G source pixels of one source row and writes the 2 x 2G destination pixels they own: 3 rows x (G+2)
// pixels in (word loads), 2 x 2*G*CN bytes out (64-bit stores) -- the kernel is bound by its 4x larger output.
template <int CN, int G>
__global__ void __launch_bounds__(128) k_resize_up2x_u8(const ResizeArgs a) {
  constexpr int IN_B = G * CN, NW = IN_B / 4 + 2;  // main words + one halo word on each side
  static_assert(IN_B % 4 == 0, "whole words");
  const int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of G source pixels
  const int j = blockIdx.y;                             // source row
  const int groups = a.scols / G;
  if (g >= groups) return;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const int jr[3] = {max(j - 1, 0), j, min(j + 1, a.srows - 1)};
  int cv[3][G + 2][CN];  // [row][pixel -1 .. G][channel]
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t *row = (const uint32_t *)(src + (size_t)jr[r] * a.sstep + (size_t)g * IN_B);
    uint32_t w[NW];
    w[0] = g > 0 ? __ldg(row - 1) : 0u;
#pragma unroll
    for (int k = 0; k < IN_B / 4; ++k) w[k + 1] = __ldg(row + k);
    w[NW - 1] = g + 1 < groups ? __ldg(row + IN_B / 4) : 0u;
#pragma unroll
    for (int p = -1; p <= G; ++p)
#pragma unroll
      for (int ch = 0; ch < CN; ++ch) cv[r][p + 1][ch] = (int)byte_at(w, 4 + p * CN + ch);
#pragma unroll
    for (int ch = 0; ch < CN; ++ch) {  // clamp the halo pixels at the ends of the row
      if (g == 0) cv[r][0][ch] = cv[r][1][ch];
      if (g + 1 == groups) cv[r][G + 1][ch] = cv[r][G][ch];
    }
  }
  uint32_t out[2][2 * IN_B / 4];
#pragma unroll
  for (int k = 0; k < 2 * IN_B / 4; ++k) out[0][k] = out[1][k] = 0;
#pragma unroll
  for (int p = 0; p < G; ++p)
#pragma unroll
    for (int e = 0; e < 2; ++e)  // destination column 2p + e: far = p-1 (even) or p+1 (odd)
#pragma unroll
      for (int ch = 0; ch < CN; ++ch) {
        const int fp = e ? p + 2 : p;  // index of the far pixel in cv (pixel + 1)
        const int hu = cv[0][fp][ch] + 3 * cv[0][p + 1][ch];
        const int hm = cv[1][fp][ch] + 3 * cv[1][p + 1][ch];
        const int hd = cv[2][fp][ch] + 3 * cv[2][p + 1][ch];
        const int m3 = (3 * hm) >> 2;
        const uint32_t v0 = (uint32_t)(((hu >> 2) + m3 + 2) >> 2);  // destination row 2j
        const uint32_t v1 = (uint32_t)(((hd >> 2) + m3 + 2) >> 2);  // destination row 2j + 1
        const int ob = (2 * p + e) * CN + ch;
        out[0][ob >> 2] |= v0 << ((ob & 3) * 8);
        out[1][ob >> 2] |= v1 << ((ob & 3) * 8);
      }
  uint8_t *d0 = a.dst + (size_t)blockIdx.z * a.dfs + (size_t)(2 * j) * a.dstep + (size_t)g * 2 * IN_B;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    uint2 *d = (uint2 *)(d0 + (size_t)rr * a.dstep);
#pragma unroll
    for (int k = 0; k < IN_B / 4; ++k) d[k] = make_uint2(out[rr][2 * k], out[rr][2 * k + 1]);
  }
}

template <int CN, int G>
static int launch_resize_up2x(const ResizeArgs &a, int n, cudaStream_t s) {
  dim3 grid(ceil_div(a.scols / G, 128), a.srows, n);
  k_resize_up2x_u8<CN, G><<<grid, 128, 0, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

int launch_resize(Ctx *c, const DBatch &src, const DBatch &dst, cudaStream_t s) {
  if (dst.v.rows == 0 || dst.v.cols == 0 || src.n == 0) return RCV_OK;
  if (src.v.rows == 0 || src.v.cols == 0) return fail(RCV_ERR_SIZE, "resize from an empty image");
  if (dst.v.rows > 65535 || src.n > 65535) return fail(RCV_ERR_UNSUPPORTED, "image too tall / batch too large");
  ResizeArgs a{src.v.data, src.v.step, src.frame_stride, dst.v.data, dst.v.step, dst.frame_stride,
               src.v.rows, src.v.cols, dst.v.rows, dst.v.cols, src.v.cn, nullptr, nullptr};
  const bool al = ((((uintptr_t)src.v.data | src.v.step | src.frame_stride) & 15) == 0) &&
                  ((((uintptr_t)dst.v.data | dst.v.step | dst.frame_stride) & 3) == 0);
  if (src.v.depth == RCV_U8 && src.v.cn == 3 && src.v.rows == 4 * dst.v.rows && src.v.cols == 4 * dst.v.cols &&
      (dst.v.cols & 3) == 0 && al && opt_get("resize.force_generic", 0) == 0) {
    dim3 grid(ceil_div(dst.v.cols >> 2, 128), dst.v.rows, src.n);
    k_resize4x_u8c3<<<grid, 128, 0, s>>>(a);
    count_launch();
    RCV_CUDA(cudaGetLastError());
    return RCV_OK;
  }
  if (src.v.depth == RCV_U8 && src.v.rows == 2 * dst.v.rows && src.v.cols == 2 * dst.v.cols && al &&
      opt_get("resize.force_generic", 0) == 0) {
    if (src.v.cn == 3 && (dst.v.cols & 7) == 0) return launch_resize2x<3, 8>(a, src.n, s);
    if (src.v.cn == 1 && (dst.v.cols & 7) == 0) return launch_resize2x<1, 8>(a, src.n, s);
    if (src.v.cn == 4 && (dst.v.cols & 3) == 0) return launch_resize2x<4, 4>(a, src.n, s);
  }
  if (src.v.depth == RCV_U8 && dst.v.rows == 2 * src.v.rows && dst.v.cols == 2 * src.v.cols && src.v.rows <= 65535 &&
      ((((uintptr_t)src.v.data | src.v.step | src.frame_stride) & 3) == 0) &&
      ((((uintptr_t)dst.v.data | dst.v.step | dst.frame_stride) & 7) == 0) && opt_get("resize.force_generic", 0) == 0) {
    if (src.v.cn == 3 && (src.v.cols & 3) == 0) return launch_resize_up2x<3, 4>(a, src.n, s);
    if (src.v.cn == 1 && (src.v.cols & 3) == 0) return launch_resize_up2x<1, 4>(a, src.n, s);
    if (src.v.cn == 4 && (src.v.cols & 1) == 0) return launch_resize_up2x<4, 2>(a, src.n, s);
  }
  // the per-column / per-row tables depend on the geometry only: rebuilt and uploaded when it changes
  void *dcols = nullptr, *drows = nullptr;
  const size_t cols8_off = ((size_t)dst.v.cols * sizeof(ResizeCol) + 15) & ~(size_t)15;
  RCV_TRY(ctx_scratch(c, SCR_TABLE_X, cols8_off + ((size_t)dst.v.cols + 3) * sizeof(int2), &dcols));
  RCV_TRY(ctx_scratch(c, SCR_TABLE_Y, (size_t)dst.v.rows * sizeof(ResizeRow), &drows));
  const int key[4] = {src.v.rows, src.v.cols, dst.v.rows, dst.v.cols};
  if (memcmp(key, c->resize_key, sizeof(key)) != 0) {
    std::vector<ResizeCol> cols;
    std::vector<ResizeRow> rows;
    resize_tables(src.v.rows, src.v.cols, dst.v.rows, dst.v.cols, cols, rows);
    RCV_CUDA(cudaMemcpyAsync(dcols, cols.data(), cols.size() * sizeof(ResizeCol), cudaMemcpyHostToDevice, s));
    std::vector<int2> cols8(cols.size() + 3);  // + 3 copies of the last entry: a group of 4 is read whole
    for (size_t i = 0; i < cols.size(); ++i)
      cols8[i] = cols[i].x1 == cols[i].x0 ? make_int2(cols[i].x0, cols[i].a0 + cols[i].a1)  // clamped: one tap takes both weights
                                          : make_int2(cols[i].x0, cols[i].a0 | (cols[i].a1 << 16));
    for (size_t i = cols.size(); i < cols8.size(); ++i) cols8[i] = cols8[cols.size() - 1];
    RCV_CUDA(cudaMemcpyAsync((uint8_t *)dcols + cols8_off, cols8.data(), cols8.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
    RCV_CUDA(cudaMemcpyAsync(drows, rows.data(), rows.size() * sizeof(ResizeRow), cudaMemcpyHostToDevice, s));
    // the vectors die at return: do not depend on how the driver stages pageable sources
    RCV_CUDA(cudaStreamSynchronize(s));
    memcpy(c->resize_key, key, sizeof(key));
  }
  a.cols = (const ResizeCol *)dcols;
  a.cols8 = (const int2 *)((const uint8_t *)dcols + cols8_off);
  a.rows = (const ResizeRow *)drows;
  // u8 BGR / BGRA with word-aligned rows on both sides: the word-load kernel (rows must be readable up to the
  // word that holds their last byte: step >= row bytes rounded up to 4)
  if (src.v.depth == RCV_U8 && src.v.cn >= 1 && src.v.cn <= 4 && opt_get("resize.force_generic", 0) == 0 &&
      opt_get("resize.byte_loads", 0) == 0 && ((((uintptr_t)src.v.data | src.v.step | src.frame_stride) & 3) == 0) &&
      ((((uintptr_t)dst.v.data | dst.v.step | dst.frame_stride) & 3) == 0) &&
      src.v.step >= (((size_t)src.v.cols * src.v.cn + 3) & ~(size_t)3)) {
    dim3 gridw(ceil_div(ceil_div(dst.v.cols, 4), 32), ceil_div(dst.v.rows, 4), src.n);
    switch (src.v.cn) {
      case 1: k_resize_u8w<1><<<gridw, 128, 0, s>>>(a); break;
      case 2: k_resize_u8w<2><<<gridw, 128, 0, s>>>(a); break;
      case 3: k_resize_u8w<3><<<gridw, 128, 0, s>>>(a); break;
      default: k_resize_u8w<4><<<gridw, 128, 0, s>>>(a); break;
    }
    count_launch();
    RCV_CUDA(cudaGetLastError());
    return RCV_OK;
  }
  dim3 grid(ceil_div(dst.v.cols, 256), dst.v.rows, src.n);
  if (src.v.depth == RCV_U8)
    k_resize_generic<uint8_t><<<grid, 256, 0, s>>>(a);
  else
    k_resize_generic<float><<<grid, 256, 0, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

// ---------------------------------------------------------------------------------------
// warpAffine
// ---------------------------------------------------------------------------------------
void rotation_matrix(double cx, double cy, double angle_deg, double scale, double M[6]) {
  double ang = angle_deg * (3.14159265358979323846 / 180.0);
  double alpha = scale * cos(ang);
  double beta = scale * sin(ang);
  M[0] = alpha;
  M[1] = beta;
  M[2] = (1 - alpha) * cx - beta * cy;
  M[3] = -beta;
  M[4] = alpha;
  M[5] = beta * cx + (1 - alpha) * cy;
}

int invert_affine(const double M[6], double iM[6]) {
  double D = M[0] * M[4] - M[1] * M[3];
  if (D == 0.0) return -1;
  D = 1.0 / D;
  double A11 = M[4] * D, A22 = M[0] * D;
  double A12 = -M[1] * D, A21 = -M[3] * D;
  double b1 = -A11 * M[2] - A12 * M[5];
  double b2 = -A21 * M[2] - A22 * M[5];
  iM[0] = A11;
  iM[1] = A12;
  iM[2] = b1;
  iM[3] = A21;
  iM[4] = A22;
  iM[5] = b2;
  return 0;
}

struct WarpArgs {
  const uint8_t *src;
  size_t sstep, sfs;
  uint8_t *dst;
  size_t dstep, dfs;
  int srows, scols, drows, dcols, cn;
  double m1, m2, m4, m5;  // row terms (f64)
  float m0, m3;           // column factors (f32)
  float border;
};

template <typename T>
__global__ void __launch_bounds__(256) k_warp_affine(const WarpArgs a) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= a.dcols || y >= a.drows) return;
  // bx = (float)(iM[1]*y + iM[2]) : f64 mul, f64 add, one rounding to f32
  const float bx = __double2float_rn(__dadd_rn(__dmul_rn(a.m1, (double)y), a.m2));
  const float by = __double2float_rn(__dadd_rn(__dmul_rn(a.m4, (double)y), a.m5));
  const float sx = fmaf(a.m0, (float)x, bx);
  const float sy = fmaf(a.m3, (float)x, by);
  const float flx = floorf(sx), fly = floorf(sy);
  const float fx = __fsub_rn(sx, flx), fy = __fsub_rn(sy, fly);
  const bool inside = flx >= -1.0f && flx < (float)a.scols && fly >= -1.0f && fly < (float)a.srows;
  const int ix = inside ? (int)flx : 0, iy = inside ? (int)fly : 0;
  const bool x0ok = inside && ix >= 0, x1ok = inside && ix + 1 < a.scols;
  const bool y0ok = inside && iy >= 0, y1ok = inside && iy + 1 < a.srows;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  const T *r0 = (const T *)(src + (size_t)(y0ok ? iy : 0) * a.sstep);
  const T *r1 = (const T *)(src + (size_t)(y1ok ? iy + 1 : 0) * a.sstep);
  T *d = (T *)(a.dst + (size_t)blockIdx.z * a.dfs + (size_t)y * a.dstep);
  for (int ch = 0; ch < a.cn; ++ch) {
    float p00 = a.border, p01 = a.border, p10 = a.border, p11 = a.border;
    if (y0ok && x0ok) p00 = (float)__ldg(r0 + ix * a.cn + ch);
    if (y0ok && x1ok) p01 = (float)__ldg(r0 + (ix + 1) * a.cn + ch);
    if (y1ok && x0ok) p10 = (float)__ldg(r1 + ix * a.cn + ch);
    if (y1ok && x1ok) p11 = (float)__ldg(r1 + (ix + 1) * a.cn + ch);
    float q0 = fmaf(fx, __fsub_rn(p01, p00), p00);
    float q1 = fmaf(fx, __fsub_rn(p11, p10), p10);
    float v = fmaf(fy, __fsub_rn(q1, q0), q0);
    if (sizeof(T) == 1) {
      int iv = __float2int_rn(v);
      d[x * a.cn + ch] = (T)min(max(iv, 0), 255);
    } else {
      d[x * a.cn + ch] = (T)v;
    }
  }
}

// ---- f32 single channel, TMA-staged source tiles -------------------------------------------
// A CTA makes one 64x32 dst tile.  The source pixels it can touch lie in the tile's rotated
// bounding box; one cp.async.bulk.tensor load brings that box (BW x BH floats, out-of-image
// parts zero-filled) into shared memory, and the four taps per pixel are gathered from shared
// memory instead of from L1/L2 (a rotated warp row touches ~9 cache lines per tap in global
// memory: the direct kernel is L1-wavefront bound, 5x off the HBM roofline).  Arithmetic and
// the in-image tests are exactly those of k_warp_affine / the oracle.
constexpr int kWarpTW = 64, kWarpThreads = 256;  // tile height TH = 32 or 64 rows (template parameter)

struct WarpTileArgs {
  WarpArgs a;
  double d0, d3;  // iM[0], iM[3] in f64 for the tile's bounding box
  int bw, bh;     // box size: bw 4-byte words per row, bh rows
};

// One pixel (all CN channels).  INTERIOR: the whole box lies inside the image, so every tap is
// in-image and inside the box -- no predicates at all.  Same operations in the same order either
// way.  T = float (CN == 1) or uint8_t (CN = 1..4); box geometry is in BYTES (bx0 is a byte offset).
template <typename T, int CN>
__device__ __forceinline__ float tap_load(uint32_t addr) {
  if (sizeof(T) == 4) return lds_f32(addr);
  return (float)lds8(addr);
}

template <typename T, int CN, bool INTERIOR>
__device__ __forceinline__ void warp_pixel(const WarpTileArgs &t, uint32_t tile, int bx0, int by0, float sx, float sy,
                                           const uint8_t *src, T *out) {
  constexpr int E = CN * (int)sizeof(T);  // bytes per pixel
  const WarpArgs &a = t.a;
  // floor through the integer: F2I.FLOOR + I2FP (ALU pipe) instead of FRND.FLOOR + F2I (two conversion-
  // pipe ops).  Identical to floorf for every finite |s| < 2^31; beyond that F2I saturates, which the
  // in-image test below rejects exactly like the oracle's float comparison does.
  int ix = __float2int_rd(sx), iy = __float2int_rd(sy);
  const float flx = __int2float_rn(ix), fly = __int2float_rn(iy);
  const float fx = __fsub_rn(sx, flx), fy = __fsub_rn(sy, fly);
  const int rowb = t.bw * 4;  // box row pitch in bytes
  bool inside = true, x0ok = true, x1ok = true, y0ok = true, y1ok = true, inbox = true;
  if (!INTERIOR) {
    inside = flx >= -1.0f && flx < (float)a.scols && fly >= -1.0f && fly < (float)a.srows;
    ix = inside ? ix : 0;
    iy = inside ? iy : 0;
    x0ok = inside && ix >= 0;
    x1ok = inside && ix + 1 < a.scols;
    y0ok = inside && iy >= 0;
    y1ok = inside && iy + 1 < a.srows;
  }
  const int cxb = ix * E - bx0, cy = iy - by0;  // byte offset of tap (0,0) inside the box
  if (!INTERIOR) inbox = cxb >= 0 && cxb + 2 * E <= rowb && (unsigned)cy < (unsigned)(t.bh - 1);
  // interior tiles: `tile` arrives with the box origin folded in (tile - by0 * rowb - bx0), so the tap address is one
  // multiply-add and one shift-add instead of four operations per pixel
  const uint32_t p = INTERIOR ? tile + (uint32_t)(iy * rowb) + (uint32_t)(ix * E) : tile + (uint32_t)(cy * rowb + cxb);
#pragma unroll
  for (int ch = 0; ch < CN; ++ch) {
    float p00, p01, p10, p11;
    if (INTERIOR) {
      p00 = tap_load<T, CN>(p + ch * sizeof(T));
      p01 = tap_load<T, CN>(p + E + ch * sizeof(T));
      p10 = tap_load<T, CN>(p + rowb + ch * sizeof(T));
      p11 = tap_load<T, CN>(p + rowb + E + ch * sizeof(T));
    } else {
      p00 = p01 = p10 = p11 = a.border;
      if (inbox) {
        const float v00 = tap_load<T, CN>(p + ch * sizeof(T)), v01 = tap_load<T, CN>(p + E + ch * sizeof(T));
        const float v10 = tap_load<T, CN>(p + rowb + ch * sizeof(T)), v11 = tap_load<T, CN>(p + rowb + E + ch * sizeof(T));
        if (y0ok && x0ok) p00 = v00;
        if (y0ok && x1ok) p01 = v01;
        if (y1ok && x0ok) p10 = v10;
        if (y1ok && x1ok) p11 = v11;
      } else if (inside) {  // cannot happen for sane boxes; keeps the result exact regardless
        const T *r0 = (const T *)(src + (size_t)(y0ok ? iy : 0) * a.sstep);
        const T *r1 = (const T *)(src + (size_t)(y1ok ? iy + 1 : 0) * a.sstep);
        if (y0ok && x0ok) p00 = (float)__ldg(r0 + ix * CN + ch);
        if (y0ok && x1ok) p01 = (float)__ldg(r0 + (ix + 1) * CN + ch);
        if (y1ok && x0ok) p10 = (float)__ldg(r1 + ix * CN + ch);
        if (y1ok && x1ok) p11 = (float)__ldg(r1 + (ix + 1) * CN + ch);
      }
    }
    const float q0 = fmaf(fx, __fsub_rn(p01, p00), p00);
    const float q1 = fmaf(fx, __fsub_rn(p11, p10), p10);
    const float v = fmaf(fy, __fsub_rn(q1, q0), q0);
    if (sizeof(T) == 1) {
      const int iv = __float2int_rn(v);
      out[ch] = (T)min(max(iv, 0), 255);
    } else {
      out[ch] = (T)v;
    }
  }
}

template <typename T, int CN, int TH>
__global__ void __launch_bounds__(kWarpThreads) k_warp_tile(const __grid_constant__ CUtensorMap tmap,
                                                            const WarpTileArgs t) {
  // no static shared memory in this kernel: the TMA destination must be 128-byte aligned and the
  // dynamic segment only starts at offset 0 when nothing static precedes it.  Layout:
  // [tile bw words x bh rows][mbarrier 8 B, box origin 8 B][row terms: TH x (bx, by) floats]
  constexpr int E = CN * (int)sizeof(T);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const WarpArgs &a = t.a;
  const int tx0 = blockIdx.x * kWarpTW, ty0 = blockIdx.y * TH;
  const uint32_t tile = smem_u32(smem_raw);
  const uint32_t barp = tile + (((uint32_t)(t.bw * t.bh * 4) + 15u) & ~15u);
  const uint32_t rowtab = barp + 16;
  if (threadIdx.x == 0) {
    // bounding box origin from the four tile corners (f64), once per tile
    const double xs[2] = {(double)tx0, (double)(tx0 + kWarpTW - 1)}, ys[2] = {(double)ty0, (double)(ty0 + TH - 1)};
    double minx = 1e300, miny = 1e300;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double sx = t.d0 * xs[i] + a.m1 * ys[j] + a.m2;
        double sy = t.d3 * xs[i] + a.m4 * ys[j] + a.m5;
        minx = fmin(minx, sx);
        miny = fmin(miny, sy);
      }
    // clamp far-out-of-image boxes so the int conversion and the TMA coordinates stay sane
    minx = fmin(fmax(minx, -1.0e6), 1.0e6 + a.scols);
    miny = fmin(fmax(miny, -1.0e6), 1.0e6 + a.srows);
    // the box starts on a 16-byte boundary of the row: TMA rejects other inner offsets
    const int ox = (((int)floor(minx) - 1) * E) & ~15, oy = (int)floor(miny) - 1;  // ox in bytes
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(barp + 8), "r"(ox), "r"(oy) : "memory");
    mbar_init(barp, 1);
    fence_mbar_init();
    fence_proxy_async();
    mbar_expect_tx(barp, (uint32_t)(t.bw * t.bh * 4));
    tma_load_3d(tile, &tmap, barp, ox >> 2, oy, (int)blockIdx.z);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + TH) {
    // row terms, once per tile row: bx = (float)(iM[1]*y + iM[2]) -- f64 mul, f64 add, one rounding
    const int r = threadIdx.x - 32;
    const double y = (double)(ty0 + r);
    const float bx = __double2float_rn(__dadd_rn(__dmul_rn(a.m1, y), a.m2));
    const float by = __double2float_rn(__dadd_rn(__dmul_rn(a.m4, y), a.m5));
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(rowtab + r * 8), "f"(bx), "f"(by) : "memory");
  }
  __syncthreads();  // barrier init, box origin and row terms visible
  int bx0, by0;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bx0), "=r"(by0) : "r"(barp + 8));
  const int lx = threadIdx.x & 63, ly0 = threadIdx.x >> 6;
  const int x = tx0 + lx;
  const float xf = (float)x;
  const uint8_t *src = a.src + (size_t)blockIdx.z * a.sfs;
  uint8_t *drow = a.dst + (size_t)blockIdx.z * a.dfs + (size_t)(ty0 + ly0) * a.dstep + (size_t)x * E;
  const bool interior = bx0 >= 0 && by0 >= 0 && bx0 + t.bw * 4 <= a.scols * E && by0 + t.bh <= a.srows;
  const bool full = tx0 + kWarpTW <= a.dcols && ty0 + TH <= a.drows;
  mbar_wait(barp, 0);
  if (interior && full) {
    const uint32_t tile0 = tile - (uint32_t)(by0 * (t.bw * 4)) - (uint32_t)bx0;  // box origin folded into the base
    // the output address as ONE global-space register pair, advanced by one 64-bit add per pixel (the compiler's own
    // form was a uniform base + a running offset = two 64-bit adds, or a 64-bit multiply-add from the tile's first row)
    unsigned long long o8 = (unsigned long long)__cvta_generic_to_global(drow);
    asm volatile("" : "+l"(o8));
    const unsigned long long step4 = 4ull * a.dstep;
#pragma unroll
    for (int k = 0; k < TH / 4; ++k) {
      float bx, by;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bx), "=f"(by) : "r"(rowtab + (ly0 + 4 * k) * 8));
      T v[CN];
      warp_pixel<T, CN, true>(t, tile0, bx0, by0, fmaf(a.m0, xf, bx), fmaf(a.m3, xf, by), src, v);
#pragma unroll
      for (int ch = 0; ch < CN; ++ch) {
        if (sizeof(T) == 4)
          asm volatile("st.global.f32 [%0], %1;" ::"l"(o8 + ch * 4), "f"((float)v[ch]) : "memory");
        else
          asm volatile("st.global.u8 [%0], %1;" ::"l"(o8 + ch), "r"((uint32_t)v[ch]) : "memory");
      }
      o8 += step4;
    }
  } else {
#pragma unroll 2
    for (int k = 0; k < TH / 4; ++k) {
      const int y = ty0 + ly0 + 4 * k;
      float bx, by;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bx), "=f"(by) : "r"(rowtab + (ly0 + 4 * k) * 8));
      T v[CN];
      warp_pixel<T, CN, false>(t, tile, bx0, by0, fmaf(a.m0, xf, bx), fmaf(a.m3, xf, by), src, v);
      if (x < a.dcols && y < a.drows) {
        T *o = (T *)(drow + (size_t)(4 * k) * a.dstep);
#pragma unroll
        for (int ch = 0; ch < CN; ++ch) o[ch] = v[ch];
      }
    }
  }
}

// Simulates the tap addresses of a few warps (32 consecutive dst x) and returns the pitch, among
// min_words rounded up to 4 plus {0, 4, .., 28}, with the fewest shared-memory wavefronts per load.
static int pick_box_pitch(int min_words, int E, double m0, double m3) {
  // the answer depends on the matrix only: remember the last one (a capture loop warps with one matrix)
  static thread_local struct { double m0, m3; int e, mw, bw; } last = {0, 0, 0, 0, 0};
  if (last.bw && last.m0 == m0 && last.m3 == m3 && last.e == E && last.mw == min_words) return last.bw;
  const int base = (min_words + 3) & ~3;
  int best = base;
  long best_cost = -1;
  for (int add = 0; add < 32; add += 4) {
    const int rowb = (base + add) * 4;
    long cost = 0;
    for (int trial = 0; trial < 12; ++trial) {
      int word[32];
      for (int l = 0; l < 32; ++l) {
        const double sx = m0 * (37 * trial + l) + 0.37 * trial, sy = m3 * (37 * trial + l) + 0.61 * trial;
        const long addr = (long)floor(sy) * rowb + (long)floor(sx) * E;
        word[l] = (int)(addr >> 2);  // arithmetic shift: floor for negatives
      }
      int worst = 1;
      for (int bank = 0; bank < 32; ++bank) {
        int distinct = 0, seen[32];
        for (int l = 0; l < 32; ++l) {
          if ((word[l] & 31) != bank) continue;
          bool dup = false;
          for (int k = 0; k < distinct; ++k) dup |= seen[k] == word[l];
          if (!dup) seen[distinct++] = word[l];
        }
        if (distinct > worst) worst = distinct;
      }
      cost += worst;
    }
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = base + add;
    }
  }
  last.m0 = m0;
  last.m3 = m3;
  last.e = E;
  last.mw = min_words;
  last.bw = best;
  return best;
}

template <typename T, int CN, int TH>
static int launch_warp_tile(const CUtensorMap &tmap, const WarpTileArgs &t, dim3 grid, size_t smem, cudaStream_t s) {
  static size_t attr_smem[16] = {};  // per instantiation, per device: largest size set so far
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_smem[dev & 15] < smem) {
    RCV_CUDA(cudaFuncSetAttribute(k_warp_tile<T, CN, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[dev & 15] = smem;
  }
  k_warp_tile<T, CN, TH><<<grid, kWarpThreads, smem, s>>>(tmap, t);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

// Tries the TMA-staged kernel with TH-row tiles.  Returns RCV_ERR_UNSUPPORTED (without setting the error
// string) when the tile's bounding box does not fit: the caller then tries the next shape.
template <int TH>
static int try_warp_tile(const DBatch &src, const WarpArgs &a, const double iM[6], size_t smem_cap, cudaStream_t s) {
  const int E = (int)src.v.elem() * src.v.cn;  // bytes per pixel
  const double dxw = fabs(iM[0]) * (kWarpTW - 1) + fabs(iM[1]) * (TH - 1);
  const double dyh = fabs(iM[3]) * (kWarpTW - 1) + fabs(iM[4]) * (TH - 1);
  if (!(dxw < 250.0 && dyh < 250.0)) return RCV_ERR_UNSUPPORTED;
  WarpTileArgs t;
  t.a = a;
  t.d0 = iM[0];
  t.d3 = iM[3];
  // box width in 4-byte words: (ceil(dxw) + 4) pixels, + 15 bytes because the origin is floored to a
  // 16-byte boundary; a multiple of 4 words (TMA inner box = multiple of 16 bytes).  Among the 8
  // residues mod 32 the one with the fewest shared-memory bank conflicts for THIS matrix is taken:
  // a warp's 32 taps walk a line of slope (m0, m3) through the box (at 90 degrees a pitch that is a
  // multiple of 32 words would be a 32-way conflict; at 15 degrees it is the conflict-free one).
  t.bw = pick_box_pitch((((int)ceil(dxw) + 4) * E + 15 + 3) / 4, E, iM[0], iM[3]);
  t.bh = (int)ceil(dyh) + 4;
  const size_t smem = (((size_t)t.bw * t.bh * 4 + 15) & ~(size_t)15) + 16 + TH * 8;  // tile + mbarrier + row terms
  dim3 grid(ceil_div(a.dcols, kWarpTW), ceil_div(a.drows, TH), src.n);
  if (!(t.bw <= 256 && t.bh <= 256 && smem <= smem_cap && grid.y <= 65535)) return RCV_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  RCV_TRY(make_tmap_rows_u32(&tmap, src.v.data, src.v.row_bytes(), src.v.rows, src.v.step, src.n, src.frame_stride,
                             t.bw, t.bh));
  if (src.v.depth == RCV_F32) return launch_warp_tile<float, 1, TH>(tmap, t, grid, smem, s);
  switch (src.v.cn) {
    case 1: return launch_warp_tile<uint8_t, 1, TH>(tmap, t, grid, smem, s);
    case 2: return launch_warp_tile<uint8_t, 2, TH>(tmap, t, grid, smem, s);
    case 3: return launch_warp_tile<uint8_t, 3, TH>(tmap, t, grid, smem, s);
    case 4: return launch_warp_tile<uint8_t, 4, TH>(tmap, t, grid, smem, s);
  }
  return RCV_ERR_UNSUPPORTED;
}

int launch_warp_affine(Ctx *c, const DBatch &src, const DBatch &dst, const double iM[6], double border,
                       cudaStream_t s) {
  if (dst.v.rows == 0 || dst.v.cols == 0 || src.n == 0) return RCV_OK;
  if (src.n > 65535) return fail(RCV_ERR_UNSUPPORTED, "batch too large");
  WarpArgs a;
  a.src = src.v.data;
  a.sstep = src.v.step;
  a.sfs = src.frame_stride;
  a.dst = dst.v.data;
  a.dstep = dst.v.step;
  a.dfs = dst.frame_stride;
  a.srows = src.v.rows;
  a.scols = src.v.cols;
  a.drows = dst.v.rows;
  a.dcols = dst.v.cols;
  a.cn = src.v.cn;
  a.m0 = (float)iM[0];
  a.m1 = iM[1];
  a.m2 = iM[2];
  a.m3 = (float)iM[3];
  a.m4 = iM[4];
  a.m5 = iM[5];
  a.border = src.v.depth == RCV_U8 ? (float)(int)border : (float)border;

  // TMA-staged path: f32 C1 or u8 C1..C4, 16-byte aligned source, bounding box small enough for smem.
  // 64x64 (else 64x48) tiles when the box stays under 36 KB: the per-tile prologue -- corner arithmetic in
  // f64, barrier set-up, row terms, ~100 instructions per thread -- is amortised over 16 (12) pixels per
  // thread instead of 8, and the box over-read shrinks (15 degrees: 1.7x vs 2.0x); else 64x32 tiles up to 96 KB.
  const bool tile_type = (src.v.depth == RCV_F32 && src.v.cn == 1) || src.v.depth == RCV_U8;
  if (tile_type && src.v.rows > 0 && src.v.cols > 0 &&
      ((((uintptr_t)src.v.data | src.v.step | src.frame_stride) & 15) == 0) && opt_get("warp.force_generic", 0) == 0) {
    int rc = RCV_ERR_UNSUPPORTED;
    const int64_t th = opt_get("warp.tile_rows", 0);  // 0 = automatic; 32 / 48 / 64 force one shape
    const size_t big = 96 * 1024, cap = 36 * 1024;    // 36 KB boxes: 6 CTAs (48 warps) per SM
    // measured (profiles/README.md): f32 C1 15/33/90 degrees 64 rows best (0.82 / 0.78 / 0.54 of the roofline vs
    // 0.73 / 0.71 / 0.51 at 32); u8 BGR gains only at 48 (the 64-row loop of 3 channels x 16 pixels is too long)
    const bool f32 = src.v.depth == RCV_F32;
    if ((th == 0 && f32 && a.drows >= 64) || th == 64) rc = try_warp_tile<64>(src, a, iM, th ? big : cap, s);
    if (rc == RCV_ERR_UNSUPPORTED && ((th == 0 && a.drows >= 48) || th == 48)) rc = try_warp_tile<48>(src, a, iM, th ? big : cap, s);
    if (rc == RCV_ERR_UNSUPPORTED) rc = try_warp_tile<32>(src, a, iM, big, s);
    if (rc != RCV_ERR_UNSUPPORTED) return rc;
  }

  dim3 block(32, 8, 1);
  dim3 grid(ceil_div(a.dcols, 32), ceil_div(a.drows, 8), src.n);
  if (grid.y > 65535) return fail(RCV_ERR_UNSUPPORTED, "image too tall");
  if (src.v.depth == RCV_U8)
    k_warp_affine<uint8_t><<<grid, block, 0, s>>>(a);
  else
    k_warp_affine<float><<<grid, block, 0, s>>>(a);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

}  // namespace rcv
