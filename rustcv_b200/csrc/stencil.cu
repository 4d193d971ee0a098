// stencil.cu -- TMA strip-pipeline stencil kernels for sm_100a:
//   k_strip<Gauss5Op<CN>>   5x5 binomial GaussianBlur on u8 (the BASELINE.json metric kernel)
//   k_strip<Sobel3Op>       Sobel 3x3 + gradient magnitude on f32
//
// The reference has no such ops (rustcv/src/imgproc/mod.rs:1-4 is drawing only); the
// semantics are the oracle's (oracle/rcv_oracle.c: orc_sepfilter_u8_q8, orc_sobel3_f32),
// which are OpenCV's (README.md:19,30 "OpenCV parity").
//
// Design:
//   * The image is a byte tensor [frames][rows][step]; a work item is one 480-byte-wide
//     column strip x one band of rows of one frame.
//   * One WARP per work item, each warp an independent producer/consumer pipeline:
//     lane 0 issues cp.async.bulk.tensor (TMA) loads of R-row x 512-byte boxes into the
//     warp's private ring of S shared-memory stages, each guarded by an mbarrier; the
//     warp waits on the mbarrier, reads its rows with conflict-free 128-bit LDS and
//     refills the stage.  512 B = 32 lanes x 16 B; lanes 0 and 31 are halo lanes, so
//     the strip's 480 output bytes are written by lanes 1..30 as 128-bit STG.
//   * The warp marches DOWN the strip keeping the last 2*HV rows in registers, so every
//     source row is read from shared memory exactly once and the vertical pass needs no
//     re-reads; the horizontal pass takes its neighbours by warp shuffle.
//   * u8 arithmetic is SIMD-in-register: two samples per 32-bit register as 16-bit lanes
//     (vertical sums <= 4080, final sums <= 65408 fit exactly), PRMT for the stride-CN
//     byte gathers.
//   * BORDER_REFLECT_101 is produced by patching the landed tile in shared memory
//     (TMA out-of-bounds fill is zeros), only in edge strips/bands.
#include "rcv_internal.cuh"
#include "tma_ptx.cuh"

namespace rcv {

// ---------------------------------------------------------------------------------------
// strip geometry
// ---------------------------------------------------------------------------------------
constexpr int kLaneBytes = 16;
constexpr int kTileBytes = 512;  // one tile row: 32 lanes x 16 B
constexpr int kOutBytes = 480;   // lanes 1..30

struct StripOut {
  uint8_t *data;  // NULL = not wanted
  size_t step, fs;
};

struct StripParams {
  StripOut out[3];
  int rows;
  int row_bytes;  // cols * cn * elemsize
  int strips, bands, band_rows;
  int n_frames;
  int vec_store;  // all outputs 16-byte aligned (base, step, frame stride)
  long long total_items;
  unsigned long long *next_item;  // dynamic scheduler: items beyond the first round (NULL = static stride)
};

// ---------------------------------------------------------------------------------------
// Op: 5x5 binomial Gaussian on u8, CN interleaved channels.
//   out = (sum_ij k_i k_j p + 128) >> 8, k = {1,4,6,4,1}   (oracle: orc_sepfilter_u8_q8
//   with Q8 taps {16,64,96,64,16}: (sum ky kx p + 32768) >> 16 is the same number)
// ---------------------------------------------------------------------------------------
// Integer ops pinned with inline PTX so that NVVM cannot re-associate the sums (it turns
// the 4-op forms below into 5): ptxas still picks the pipe (IADD3 / IMAD.IADD / LEA).
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("{\n\t.reg .u32 t;\n\tadd.u32 t, %1, %2;\n\tadd.u32 %0, t, %3;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <int M>
__device__ __forceinline__ uint32_t madc(uint32_t a, uint32_t c) {  // a * M + c
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "n"(M), "r"(c));
  return d;
}

template <int CN>
struct Gauss5Op {
  static constexpr int HV = 2;   // rows of vertical halo
  static constexpr int P = 2;    // pixels of horizontal halo
  static constexpr int E = CN;   // bytes per pixel
  static constexpr int NOUT = 1;
  uint32_t win[4][8];  // last 4 rows, unpacked: [2w] = bytes 0,2 of word w; [2w+1] = bytes 1,3

  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 8; ++h) win[j][h] = 0;
  }

  // J = (feed index) & 3, compile time: win[J] holds the oldest row.
  // FAST: interior rows of an aligned, non-ragged strip -- always emits, lanes store 16 B or nothing.
  template <int J, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    uint32_t in[8];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      in[2 * k] = __byte_perm(w[k], 0, 0x4240);      // (b0, b2) as 16-bit lanes
      in[2 * k + 1] = __byte_perm(w[k], 0, 0x4341);  // (b1, b3)
    }
    // vertical: V = r0 + 4 r1 + 6 r2 + 4 r3 + r4 (+8 per lane: with horizontal taps summing
    // to 16 that is the final "+128" rounding term).  V <= 4088 per 16-bit lane.
    uint32_t V[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const uint32_t r0 = win[J][h], r1 = win[(J + 1) & 3][h], r2 = win[(J + 2) & 3][h], r3 = win[(J + 3) & 3][h];
      const uint32_t a = add3(r0, in[h], 0x00080008u);
      const uint32_t b = add2(r1, r3);
      V[h] = madc<6>(r2, madc<4>(b, a));
      win[J][h] = in[h];
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

// ---------------------------------------------------------------------------------------
// Op: Sobel 3x3 on single-channel f32 + magnitude.  Operation order is the oracle's
// (orc_sobel3_f32): every op a single rounded f32 op, no fma.
// out[0] = mag; ALL = true adds out[1] = gx, out[2] = gy (each optional).
// ---------------------------------------------------------------------------------------
template <bool ALL>
struct Sobel3Op {
  static constexpr int HV = 1;
  static constexpr int P = 1;
  static constexpr int E = 4;
  static constexpr int NOUT = ALL ? 3 : 1;
  float win[2][4];

  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 4; ++h) win[j][h] = 0.0f;
  }

  template <int J, bool FAST>
  __device__ __forceinline__ void feed(const uint4 &q, bool emit, uint8_t *const *outp, int nvalid, bool vec) {
    constexpr int JJ = J & 1;
    const float pp[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
    float s[6], d[6];  // columns -1..4 at index +1
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float pm = win[JJ][c], p0 = win[JJ ^ 1][c];
      float t = __fadd_rn(pm, pp[c]);
      float u = __fmul_rn(2.0f, p0);
      s[c + 1] = __fadd_rn(t, u);
      d[c + 1] = __fsub_rn(pp[c], pm);
      win[JJ][c] = pp[c];
    }
    if (!FAST && !emit) return;
    s[0] = __shfl_up_sync(0xffffffffu, s[4], 1);
    d[0] = __shfl_up_sync(0xffffffffu, d[4], 1);
    s[5] = __shfl_down_sync(0xffffffffu, s[1], 1);
    d[5] = __shfl_down_sync(0xffffffffu, d[1], 1);
    float gx[4], gy[4], mg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      gx[c] = __fsub_rn(s[c + 2], s[c]);
      float t = __fadd_rn(d[c], d[c + 2]);
      float u = __fmul_rn(2.0f, d[c + 1]);
      gy[c] = __fadd_rn(t, u);
      float xx = __fmul_rn(gx[c], gx[c]);
      float yy = __fmul_rn(gy[c], gy[c]);
      mg[c] = __fsqrt_rn(__fadd_rn(xx, yy));
    }
    store4<FAST>(outp[0], mg, nvalid, vec);
    if (ALL) {
      store4<FAST>(outp[1], gx, nvalid, vec);
      store4<FAST>(outp[2], gy, nvalid, vec);
    }
  }

  template <bool FAST>
  static __device__ __forceinline__ void store4(uint8_t *op, const float (&v)[4], int nvalid, bool vec) {
    float *o = (float *)op;
    if (ALL && !o) return;
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

// ---------------------------------------------------------------------------------------
// the strip pipeline
// ---------------------------------------------------------------------------------------
template <class Op, int R, int S, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_strip(const __grid_constant__ CUtensorMap tmap, const StripParams p) {
  static_assert(R % 4 == 0 && R >= 2 * Op::HV + 1 && R <= 32, "chunk rows");
  static_assert(Op::E * (Op::P + 1) <= 16, "horizontal halo must fit the 16-byte halo lanes");
  constexpr int HV = Op::HV, P = Op::P, E = Op::E;
  constexpr uint32_t kStageBytes = R * kTileBytes;
  extern __shared__ __align__(128) uint8_t smem_raw[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = smem_u32(smem_raw) + (uint32_t)warp * S * kStageBytes;
  const uint32_t bars = smem_u32(smem_raw) + (uint32_t)NW * S * kStageBytes + (uint32_t)warp * S * 8;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(bars + s * 8, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncwarp();

  uint32_t phase = 0;  // bit s = parity the next wait on stage s must see
  const long long total_warps = (long long)gridDim.x * NW;
  Op op;

  long long item = (long long)blockIdx.x * NW + warp;
  while (item < p.total_items) {
    // claim the next item now; the atomic's latency hides behind this item's work
    unsigned long long claimed = 0;
    if (p.next_item != nullptr && lane == 0) claimed = atomicAdd(p.next_item, 1ULL);
    const int strip = (int)(item % p.strips);
    const long long t = item / p.strips;
    const int band = (int)(t % p.bands);
    const int frame = (int)(t / p.bands);
    const int x0 = strip * kOutBytes;
    const int y0 = band * p.band_rows;
    const int y1 = min(y0 + p.band_rows, p.rows);
    const int ys = y0 - HV;                // first row fed
    const int n_feed = (y1 - y0) + 2 * HV;  // rows fed: ys .. y1+HV-1
    const int n_chunks = (n_feed + R - 1) / R;
    const int cx = (x0 - kLaneBytes) >> 2;  // word coordinate of the tile (may be -4)
    const bool left_edge = (x0 == 0);
    const bool right_edge = (p.row_bytes < x0 + kOutBytes + kLaneBytes);
    const int xr = kLaneBytes + (p.row_bytes - x0);  // tile byte offset of the first byte past the row
    const bool top = (ys < 0);
    const bool bottom = (y1 + HV > p.rows);  // the fed rows run past the last image row
    const bool fast_strip = p.vec_store != 0 && !right_edge;  // full 16-byte stores in lanes 1..30

    // this lane's slice of the outputs
    const int xl = x0 + (lane - 1) * kLaneBytes;
    int nvalid = 0;
    if (lane >= 1 && lane <= 30) nvalid = min(max(p.row_bytes - xl, 0), kLaneBytes);
    // output pointers of the next row to emit (row y0), advanced by one step per emitted row
    uint8_t *optr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      optr[k] = (k < Op::NOUT && p.out[k].data)
                    ? p.out[k].data + (size_t)frame * p.out[k].fs + (size_t)y0 * p.out[k].step + xl
                    : nullptr;

    auto issue = [&](int c) {
      const uint32_t st = (uint32_t)(c % S);
      fence_proxy_async();
      mbar_expect_tx(bars + st * 8, kStageBytes);
      tma_load_3d(tiles + st * kStageBytes, &tmap, bars + st * 8, cx, ys + c * R, frame);
    };

    if (lane == 0) {
      for (int c = 0; c < S && c < n_chunks; ++c) issue(c);
    }
    op.reset();

    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t st = (uint32_t)(c % S);
      const uint32_t tile = tiles + st * kStageBytes;
      mbar_wait(bars + st * 8, (phase >> st) & 1u);
      phase ^= 1u << st;

      // ---- BORDER_REFLECT_101 patches (edge strips / bands only; warp-uniform branches) ----
      if (left_edge || right_edge) {
        if (lane < R) {
          const uint32_t row = tile + lane * kTileBytes;
          if (left_edge) {
#pragma unroll
            for (int k = 1; k <= P; ++k)
#pragma unroll
              for (int b = 0; b < E; ++b) sts8(row + kLaneBytes - k * E + b, lds8(row + kLaneBytes + k * E + b));
          }
          if (right_edge) {
#pragma unroll
            for (int k = 1; k <= P; ++k)
#pragma unroll
              for (int b = 0; b < E; ++b) {
                const int dsto = xr + (k - 1) * E + b;
                if (dsto < kTileBytes) sts8(row + dsto, lds8(row + xr - (k + 1) * E + b));
              }
          }
        }
        __syncwarp();
      }
      if (top && c == 0) {
        // global row -k (slot HV-k) <- row k (slot HV+k)
#pragma unroll
        for (int k = 1; k <= HV; ++k)
          sts128(tile + (HV - k) * kTileBytes + lane * kLaneBytes, lds128(tile + (HV + k) * kTileBytes + lane * kLaneBytes));
      }
      if (bottom) {
        // global row rows-1+k <- row rows-1-k; the source is in this chunk or the previous one
#pragma unroll
        for (int k = 1; k <= HV; ++k) {
          const int fi = p.rows - 1 + k - ys;  // feed index of the reflected row
          if (fi / R == c) {
            const int fs_ = fi - 2 * k;
            const uint32_t src_tile = tiles + (uint32_t)((fs_ / R) % S) * kStageBytes;
            sts128(tile + (fi % R) * kTileBytes + lane * kLaneBytes,
                   lds128(src_tile + (fs_ % R) * kTileBytes + lane * kLaneBytes));
          }
        }
      }
      __syncwarp();
      // every lane is past chunk c-1: refill its stage
      if (lane == 0 && c >= 1 && c - 1 + S < n_chunks) issue(c - 1 + S);

      // ---- rows of this chunk ----
      if (fast_strip && c * R >= 2 * HV && c * R + R <= n_feed) {
        // interior chunk of an aligned, non-ragged strip: every row is fed and emitted, no per-row tests
        const uint32_t rowaddr = tile + lane * kLaneBytes;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const uint4 q = lds128(rowaddr + j * kTileBytes);
          if ((j & 3) == 0) op.template feed<0, true>(q, true, optr, nvalid, true);
          if ((j & 3) == 1) op.template feed<1, true>(q, true, optr, nvalid, true);
          if ((j & 3) == 2) op.template feed<2, true>(q, true, optr, nvalid, true);
          if ((j & 3) == 3) op.template feed<3, true>(q, true, optr, nvalid, true);
#pragma unroll
          for (int k = 0; k < Op::NOUT; ++k)
            if (Op::NOUT == 1 || optr[k]) optr[k] += p.out[k].step;
        }
        continue;
      }
#pragma unroll 1
      for (int g = 0; g < R / 4; ++g) {
        const int fi0 = c * R + g * 4;
        if (fi0 >= n_feed) break;
        const uint32_t rowaddr = tile + (uint32_t)(g * 4) * kTileBytes + lane * kLaneBytes;
        // feed index fi produces output row y0 + fi - 2*HV
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int fi = fi0 + j;
          if (fi < n_feed) {
            const uint4 q = lds128(rowaddr + j * kTileBytes);
            const bool emit = fi >= 2 * HV;
            if (j == 0) op.template feed<0, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 1) op.template feed<1, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 2) op.template feed<2, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 3) op.template feed<3, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (emit) {
#pragma unroll
              for (int k = 0; k < Op::NOUT; ++k)
                if (optr[k]) optr[k] += p.out[k].step;
            }
          }
        }
      }
    }
    __syncwarp();  // all lanes done with the ring before the next item's prologue refills it
    if (p.next_item != nullptr)
      item = total_warps + (long long)__shfl_sync(0xffffffffu, claimed, 0);
    else
      item += total_warps;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr int kR = 8, kS = 3, kNW = 16;

static bool aligned16(const DBatch &b) {
  return b.v.data && (((uintptr_t)b.v.data | b.v.step | b.frame_stride) & 15) == 0;
}

bool strip_path_ok(const DBatch &src, int min_rows, int min_cols) {
  return aligned16(src) && src.v.rows >= min_rows && src.v.cols >= min_cols;
}

// Picks the band height.  Measured on B200 (profiles/README.md, sweep r1d): short bands win
// although every band re-feeds 2*hv warm-up rows -- the work-claim scheduler balances better
// with many small items, and the rows in flight across all resident warps then span a few
// tens of MB (TLB reach, L2-resident halo rows) instead of hundreds.  36 rows (40 fed rows =
// 5 chunks of R) measured best for the 5x5 Gaussian: 8.04 us per 4K frame vs 8.64 at 60 rows
// and 10.9 at 244.  Tiny jobs use shorter bands to occupy more warps.
static int pick_band_rows(Ctx *c, const char *optname, int rows, int strips, int n, int hv) {
  int64_t forced = opt_get(optname, 0);
  if (forced > 0) return (int)forced;
  const long long warps = (long long)ctx_sm_count(c) * kNW;
  int br = 5 * kR - 2 * hv;
  const long long items = (long long)strips * n * ((rows + br - 1) / br);
  if (items < warps) br = 4 * kR - 2 * hv;
  return br;
}

template <class Op>
static int launch_strip(Ctx *c, const DBatch &src, const DBatch *outs, int nout, const char *band_opt,
                        cudaStream_t s) {
  CUtensorMap tmap;
  RCV_TRY(make_tmap_rows_u32(&tmap, src.v.data, src.v.row_bytes(), src.v.rows, src.v.step, src.n, src.frame_stride,
                             kTileBytes / 4, kR));
  StripParams p = {};
  p.vec_store = 1;
  for (int k = 0; k < 3; ++k) {
    if (k < nout && outs[k].v.data) {
      p.out[k] = StripOut{outs[k].v.data, outs[k].v.step, outs[k].frame_stride};
      if (!aligned16(outs[k])) p.vec_store = 0;
    } else {
      p.out[k] = StripOut{nullptr, 0, 0};
    }
  }
  p.rows = src.v.rows;
  p.row_bytes = (int)src.v.row_bytes();
  p.strips = ceil_div(p.row_bytes, kOutBytes);
  p.band_rows = pick_band_rows(c, band_opt, p.rows, p.strips, src.n, Op::HV);
  p.bands = ceil_div(p.rows, p.band_rows);
  p.n_frames = src.n;
  p.total_items = (long long)p.strips * p.bands * src.n;
  p.next_item = nullptr;

  auto kern = k_strip<Op, kR, kS, kNW>;
  const int smem = kNW * kS * kR * kTileBytes + kNW * kS * 8;
  RCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long blocks = (p.total_items + kNW - 1) / kNW;
  int64_t grid_opt = opt_get("strip.grid", 0);
  int grid = (int)(blocks < ctx_sm_count(c) ? blocks : ctx_sm_count(c));
  if (grid_opt > 0) grid = (int)grid_opt;
  if (opt_get("strip.dynamic", 1) != 0 && p.total_items > (long long)grid * kNW) {
    void *ctr = nullptr;
    RCV_TRY(ctx_scratch(c, SCR_COUNTER, sizeof(unsigned long long), &ctr));
    RCV_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));
    p.next_item = (unsigned long long *)ctr;
  }
  kern<<<grid, kNW * 32, smem, s>>>(tmap, p);
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

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

int launch_sobel_strip(Ctx *c, const DBatch &src, const DBatch &mag, const DBatch &gx, const DBatch &gy,
                       cudaStream_t s) {
  if (!strip_path_ok(src, 8, 8) || src.v.depth != RCV_F32 || src.v.cn != 1) return RCV_ERR_UNSUPPORTED;
  DBatch outs[3] = {mag, gx, gy};
  if (!mag.v.data) return RCV_ERR_UNSUPPORTED;  // gx/gy without the magnitude: generic kernel
  if (!gx.v.data && !gy.v.data) return launch_strip<Sobel3Op<false>>(c, src, outs, 1, "sobel.band_rows", s);
  return launch_strip<Sobel3Op<true>>(c, src, outs, 3, "sobel.band_rows", s);
}

}  // namespace rcv
