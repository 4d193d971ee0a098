// strip_pipeline.cuh -- the TMA strip pipeline shared by the stencil / fused-chain kernels for sm_100a:
//   k_strip<Gauss5Op<CN>>      5x5 binomial GaussianBlur on u8 (the BASELINE.json metric kernel)   strip_gauss5.cu
//   k_strip<Gauss3Op<CN>>      3x3 binomial GaussianBlur on u8                                      strip_gauss3.cu
//   k_strip<GaussQ8Op<CN,KS>>  any-sigma 3x3 / 5x5 / 7x7 GaussianBlur on u8                          strip_gaussq8*.cu
//   k_strip<GaussQ8WideOp<CN,KS>>  any-sigma 9x9 .. 15x15 GaussianBlur on u8                          strip_gaussq8_wide.cuh
//   k_strip<Sobel3Op<ALL>>     Sobel 3x3 + gradient magnitude on f32                                strip_sobel.cu
//   k_strip<SepF32Op / SepF32CnOp / SepF32WideOp>      separable filters on f32, 3 .. 15 taps       strip_f32*.cu
//   k_strip<Filter2dF32CnOp / Filter2dU8Op>            dense 3x3 / 5x5 / 7x7 filters (transposed form) strip_f32cn.cu, strip_f2d_u8.cu
//   k_strip<YuyvSobelOp>, k_strip<YuyvGauss5Op>        fused decode -> process chains               strip_yuyv_*.cu
//
// The reference has no such ops (rustcv/src/imgproc/mod.rs:1-4 is drawing only); the
// semantics are the oracle's (oracle/rcv_oracle.c: orc_sepfilter_u8_q8, orc_sobel3_f32, ...),
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
//   * The warp marches DOWN the strip keeping either the last 2*HV rows or -- transposed form -- the partial
//     sums of the 2*HV outputs still open in registers, so every source row is read from shared memory
//     exactly once and the vertical pass needs no re-reads; the horizontal pass takes its neighbours by
//     warp shuffle.
//   * u8 arithmetic is SIMD-in-register: two samples per 32-bit register as 16-bit lanes
//     (vertical sums <= 4088, final sums <= 65408 fit exactly), PRMT for the stride-CN
//     byte gathers.
//   * BORDER_REFLECT_101 is produced by patching the landed tile in shared memory
//     (TMA out-of-bounds fill is zeros), only in edge strips/bands.
//   * A launch may be restricted to a row window (the banded host pipeline of abi.cu), and a job smaller
//     than the machine is cut so that its items fill one round (pick_band_rows).
#pragma once

#include "fastdiv.h"
#include "rcv_internal.cuh"
#include "tma_ptx.cuh"

#include <type_traits>

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
  int row_begin, row_end;  // output rows produced by this launch (the whole image unless the batch is windowed)
  int row_bytes;  // cols * cn * elemsize
  int strips, bands, band_rows;
  int n_frames;
  int vec_store;  // all outputs 16-byte aligned (base, step, frame stride)
  long long total_items;
  uint32_t div_strips_mul, div_strips_sh;  // item / strips == umulhi(item, mul) >> sh for every item of this launch (mul = 0: divide)
  uint32_t div_bands_mul, div_bands_sh;
  unsigned long long *next_item;  // dynamic scheduler: items beyond the first round (NULL = static stride)
  unsigned long long *reset_item; // the OTHER counter of the pair: zeroed here for the next launch on this stream
  uint32_t taps_x[4], taps_y[4];  // GaussQ8Op: symmetric Q8 taps, [0] outermost .. [KS/2] centre
  uint32_t wtaps[16];             // GaussQ8WideOp (9..15 taps): x taps [0..7], y taps [8..15], same order
  float ftaps[52];                // SepF32Op: kx[0..KS) then ky[0..KS); Filter2dOp: KS*KS taps row-major, then delta
};

// ---------------------------------------------------------------------------------------
// packed 16-bit lane arithmetic shared by the u8 ops
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

// Optional op traits (absent = 1 / 0):
//   Op::OMUL, Op::ODIV  output bytes per input byte = OMUL / ODIV: a lane that owns 16 input bytes writes
//             16*OMUL/ODIV output bytes (the fused YUYV ops turn 8 pixels of 2 bytes into 8 floats or 8 BGR pixels);
//   Op::SINGLE_PATH  only the per-row-tested loop is compiled (ops whose row body is so long that a second,
//             unpredicated copy of it would push the kernel out of the 32 KB instruction cache);
//   Op::BAND_ROWS  output rows per work item when the job is large (default 40 - 2*HV);
//   Op::EDGES the op applies the horizontal border itself (on its vertical sums, in registers) and is told
//             per item where the row's edges are: op.edges(left_edge, right_edge, xr, lane);
//   Op::HOIST_WARM  the steady-state loop carries no warm-up tests: the 2*HV window-filling rows of a band's first
//             chunk run as straight-line code (warm<j> with j = the row of the band) and the loop takes over after
//             them; an op whose UNROLL does not divide 2*HV (the 5x5 binomial: the whole chunk) gets a straight-line
//             copy of the first chunk's emitting rows as well;
//   Op::UNROLL_SLOW  rows unrolled in the per-row-tested loop (default UNROLL): ops without a rotating register
//             window can keep that rarely used loop small;
//   Op::MACRO horizontal border elements are macro-pixels reflected INCLUDING the edge element
//             (element -1 <- element 0, element n <- element n-1) instead of REFLECT_101.
//   Op::HALO_LANES  halo lanes per side (default 1): an op whose horizontal reach P*E exceeds 16 bytes (multi-channel
//             f32) gives up more lanes -- 2 halo lanes per side leave 448 output bytes per 512-byte tile row.
template <class Op, class = void>
struct OpOmul { static constexpr int value = 1; };
template <class Op>
struct OpOmul<Op, std::void_t<decltype(Op::OMUL)>> { static constexpr int value = Op::OMUL; };
template <class Op, class = void>
struct OpOdiv { static constexpr int value = 1; };
template <class Op>
struct OpOdiv<Op, std::void_t<decltype(Op::ODIV)>> { static constexpr int value = Op::ODIV; };
template <class Op, class = void>
struct OpEdges { static constexpr bool value = false; };
template <class Op>
struct OpEdges<Op, std::void_t<decltype(Op::EDGES)>> { static constexpr bool value = Op::EDGES; };
template <class Op, class = void>
struct OpSinglePath { static constexpr bool value = false; };
template <class Op>
struct OpSinglePath<Op, std::void_t<decltype(Op::SINGLE_PATH)>> { static constexpr bool value = Op::SINGLE_PATH; };
template <class Op, class = void>
struct OpBandRows { static constexpr int value = 5 * 8 - 2 * Op::HV; };  // 40 fed rows = 5 chunks
template <class Op>
struct OpBandRows<Op, std::void_t<decltype(Op::BAND_ROWS)>> { static constexpr int value = Op::BAND_ROWS; };
template <class Op, class = void>
struct OpHaloLanes {  // Op::HALO_LANES: 16-byte halo lanes on EACH side of the strip
  static constexpr int value = 1;
  static constexpr bool declared = false;
};
template <class Op>
struct OpHaloLanes<Op, std::void_t<decltype(Op::HALO_LANES)>> {
  static constexpr int value = Op::HALO_LANES;
  static constexpr bool declared = true;
};
template <class Op, class = void>
struct OpHoistWarm { static constexpr bool value = false; };
template <class Op>
struct OpHoistWarm<Op, std::void_t<decltype(Op::HOIST_WARM)>> { static constexpr bool value = Op::HOIST_WARM; };
template <class Op, class = void>
struct OpUnrollSlow { static constexpr int value = Op::UNROLL; };
template <class Op>
struct OpUnrollSlow<Op, std::void_t<decltype(Op::UNROLL_SLOW)>> { static constexpr int value = Op::UNROLL_SLOW; };
template <class Op, class = void>
struct OpMacro { static constexpr int value = 0; };
template <class Op>
struct OpMacro<Op, std::void_t<decltype(Op::MACRO)>> { static constexpr int value = Op::MACRO; };

// ---------------------------------------------------------------------------------------
// the strip pipeline
// ---------------------------------------------------------------------------------------
template <class Op, int R, int S, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_strip(const __grid_constant__ CUtensorMap tmap,
                                                      const __grid_constant__ StripParams p) {
  static_assert((R == 8 || R == 16) && R >= 2 * Op::HV + 1,
                "chunk rows: 8 (the ops' window rotation assumes it) or 16 for ops with more than 3 halo rows -- the border "
                "patches read at most one chunk back");
  constexpr int HL = OpHaloLanes<Op>::value;
  constexpr int kHalo = HL * kLaneBytes, kOut = kTileBytes - 2 * kHalo;  // halo bytes per side, output bytes per tile row
  static_assert(Op::E * Op::P <= kHalo && (OpHaloLanes<Op>::declared || Op::E * (Op::P + 1) <= 16),
                "horizontal halo must fit the halo lanes");
  constexpr int HV = Op::HV, P = Op::P, E = Op::E;
  constexpr int OM = OpOmul<Op>::value, OD = OpOdiv<Op>::value, RO = OpMacro<Op>::value;
  constexpr uint32_t kStageBytes = R * kTileBytes;
  extern __shared__ __align__(128) uint8_t smem_raw[];

  // the warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the ring / barrier addresses, the item
  // decode and the TMA operands on the uniform datapath (no R2UR + ELECT waterfall around UTMALDG)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t tiles = smem_u32(smem_raw) + (uint32_t)warp * S * kStageBytes;
  const uint32_t bars = smem_u32(smem_raw) + (uint32_t)NW * S * kStageBytes + (uint32_t)warp * S * 8;

  // Programmatic dependent launch: the launcher sets cudaLaunchAttributeProgrammaticStreamSerialization, so the
  // NEXT kernel of the stream may be scheduled as soon as this grid's CTAs have started (its CTAs become resident
  // as ours exit) and runs its prologue -- launch latency, parameter / tensor-map fetch, mbarrier set-up -- under
  // our tail.  Nothing below griddepcontrol.wait (global reads, stores, the work counters) happens before the
  // previous grid has completed and flushed.  Without the attribute both instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;");
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(bars + s * 8, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncwarp();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // Two work counters alternate between launches: this launch claims from one and clears the other,
  // which the next launch on the (single, in-order) library stream will use -- no memset node.
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.reset_item != nullptr) *p.reset_item = 0ULL;

  uint32_t phase = 0;  // bit s = parity the next wait on stage s must see
  bool dirty = false;  // lane 0: generic-proxy writes (border patches) may be pending in the ring
  const long long total_warps = (long long)gridDim.x * NW;
  Op op;
  op.init(p);

  long long item = (long long)blockIdx.x * NW + warp;
  while (item < p.total_items) {
    // claim the next item now; the atomic's latency hides behind this item's work
    unsigned long long claimed = 0;
    if (p.next_item != nullptr && lane == 0) claimed = atomicAdd(p.next_item, 1ULL);
    // 32-bit item arithmetic (the launcher refuses jobs of 2^31 items) with host-made reciprocals: the 64-bit
    // divisions were ~300 instructions per item (6 % of a 36-row band of the 5x5 Gaussian), 32-bit ones ~50
    const unsigned it32 = (unsigned)item;
    const unsigned t = p.div_strips_mul ? (__umulhi(it32, p.div_strips_mul) >> p.div_strips_sh) : it32 / (unsigned)p.strips;
    const int strip = (int)(it32 - t * (unsigned)p.strips);
    const unsigned fr = p.div_bands_mul ? (__umulhi(t, p.div_bands_mul) >> p.div_bands_sh) : t / (unsigned)p.bands;
    const int band = (int)(t - fr * (unsigned)p.bands);
    const int frame = (int)fr;
    const int x0 = strip * kOut;
    const int y0 = p.row_begin + band * p.band_rows;
    const int y1 = min(y0 + p.band_rows, p.row_end);
    const int ys = y0 - HV;                // first row fed
    const int n_feed = (y1 - y0) + 2 * HV;  // rows fed: ys .. y1+HV-1
    const int n_chunks = (n_feed + R - 1) / R;
    const int cx = (x0 - kHalo) >> 2;  // word coordinate of the tile (negative in the first strip)
    const bool left_edge = (x0 == 0);
    const bool right_edge = (p.row_bytes < x0 + kOut + kHalo);
    const int xr = kHalo + (p.row_bytes - x0);  // tile byte offset of the first byte past the row
    const bool top = (ys < 0);
    const bool bottom = (y1 + HV > p.rows);  // the fed rows run past the last image row
    const bool edgy = left_edge || right_edge || top || bottom;  // some chunk of this item gets patched
    // full 16-byte stores in every output lane: aligned outputs and a strip that is not cut by the row's end
    const bool fast_strip = p.vec_store != 0 && p.row_bytes >= x0 + kOut;

    // this lane's slice of the outputs
    const int xl = x0 + (lane - HL) * kLaneBytes;
    int nvalid = 0;
    if (lane >= HL && lane <= 31 - HL) nvalid = min(max(p.row_bytes - xl, 0), kLaneBytes) * OM / OD;  // in OUTPUT bytes
    // output pointers of the next row to emit (row y0), advanced by one step per emitted row
    uint8_t *optr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      optr[k] = (k < Op::NOUT && p.out[k].data)
                    ? p.out[k].data + (size_t)frame * p.out[k].fs + (size_t)y0 * p.out[k].step + (long long)xl * OM / OD
                    : nullptr;

    // Lane 0 only.  The proxy fence orders the border patches' generic-proxy writes before the TMA write that
    // recycles the stage; interior items never write the ring (their reads are complete -- consumed -- before the
    // __syncwarp that precedes every refill), so they skip it.
    auto issue = [&](int c, uint32_t st) {
      if (dirty) fence_proxy_async();
      mbar_expect_tx(bars + st * 8, kStageBytes);
      tma_load_3d(tiles + st * kStageBytes, &tmap, bars + st * 8, cx, ys + c * R, frame);
    };

    if (lane == 0) {
      dirty = dirty || edgy;  // the previous item's patches, or (from chunk 1 on) this item's
      for (int c = 0; c < S && c < n_chunks; ++c) issue(c, (uint32_t)c);
      dirty = edgy;
    }
    op.reset();
    if constexpr (OpEdges<Op>::value) op.edges(left_edge, right_edge, xr, lane);

    uint32_t st = 0;  // stage of chunk c = c % S, kept incrementally
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t tile = tiles + st * kStageBytes;
      mbar_wait(bars + st * 8, (phase >> st) & 1u);
      phase ^= 1u << st;
      const uint32_t st_prev = st == 0 ? (uint32_t)(S - 1) : st - 1;  // stage of chunk c-1 == stage of chunk c-1+S
      st = st + 1 == (uint32_t)S ? 0u : st + 1;

      // ---- BORDER_REFLECT_101 patches (edge strips / bands only; warp-uniform branches) ----
      if (edgy) {
        if (left_edge || right_edge) {
          if (lane < R) {
            const uint32_t row = tile + lane * kTileBytes;
            if (left_edge) {
#pragma unroll
              for (int k = 1; k <= P; ++k)
#pragma unroll
                for (int b = 0; b < E; ++b) sts8(row + kHalo - k * E + b, lds8(row + kHalo + (k - RO) * E + b));
            }
            if (right_edge) {
#pragma unroll
              for (int k = 1; k <= P; ++k)
#pragma unroll
                for (int b = 0; b < E; ++b) {
                  const int dsto = xr + (k - 1) * E + b;
                  if (dsto < kTileBytes) sts8(row + dsto, lds8(row + xr - (k + 1 - RO) * E + b));
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
      }
      __syncwarp();
      // every lane is past chunk c-1: refill its stage
      if (lane == 0 && c >= 1 && c - 1 + S < n_chunks) issue(c - 1 + S, st_prev);

      // ---- rows of this chunk ----
      if (!OpSinglePath<Op>::value && fast_strip && c * R + R <= n_feed) {
        // whole chunk of an aligned, full-width strip: every row is fed; every row emits except the
        // first 2*HV rows of the band (chunk 0), which only fill the window.  No per-row tests.
        // The body is unrolled Op::UNROLL rows (the op's window period) and looped R/UNROLL times:
        // unrolling all 8 rows of the wider ops would not fit the 32 KB instruction cache.
        constexpr int U = Op::UNROLL;
        static_assert(R % U == 0, "window period must divide the chunk");
        int g0 = 0;
        if constexpr (OpHoistWarm<Op>::value && (2 * HV) % U == 0) {
          if (c == 0) {  // the band's first 2*HV rows fill the window, straight-line; the steady loop takes over after them
            const uint32_t rowaddr = tile + lane * kLaneBytes;
#pragma unroll
            static_assert(2 * HV <= R && 2 * HV <= 14, "the window-filling rows lie in the band's first chunk");
            for (int j = 0; j < 2 * HV; ++j) {
              const uint4 q = lds128(rowaddr + j * kTileBytes);
              if (j == 0) op.template warm<0>(q);
              if (j == 1) op.template warm<1>(q);
              if (j == 2) op.template warm<2>(q);
              if (j == 3) op.template warm<3>(q);
              if (j == 4) op.template warm<4>(q);
              if (j == 5) op.template warm<5>(q);
              if (j == 6) op.template warm<6>(q);
              if (j == 7) op.template warm<7>(q);
              if (j == 8) op.template warm<8>(q);  // 16-row chunks: ops with more than 4 halo rows
              if (j == 9) op.template warm<9>(q);
              if (j == 10) op.template warm<10>(q);
              if (j == 11) op.template warm<11>(q);
              if (j == 12) op.template warm<12>(q);
              if (j == 13) op.template warm<13>(q);
            }
            g0 = 2 * HV / U;
          }
        }
        if constexpr (OpHoistWarm<Op>::value && (2 * HV) % U != 0) {
          if (c == 0) {  // the band's first chunk, straight-line: 2*HV rows fill the window, the rest emit
            static_assert(R == 8, "the straight-line first chunk dispatches rows 0..7");
            const uint32_t rowaddr = tile + lane * kLaneBytes;
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const uint4 q = lds128(rowaddr + j * kTileBytes);
              if (j < 2 * HV) {
                if (j == 0) op.template warm<0>(q);
                if (j == 1) op.template warm<1>(q);
                if (j == 2) op.template warm<2>(q);
                if (j == 3) op.template warm<3>(q);
                if (j == 4) op.template warm<4>(q);
                if (j == 5) op.template warm<5>(q);
                if (j == 6) op.template warm<6>(q);
                if (j == 7) op.template warm<7>(q);
              } else {
                if (j == 0) op.template feed<0, true>(q, true, optr, nvalid, true);
                if (j == 1) op.template feed<1, true>(q, true, optr, nvalid, true);
                if (j == 2) op.template feed<2, true>(q, true, optr, nvalid, true);
                if (j == 3) op.template feed<3, true>(q, true, optr, nvalid, true);
                if (j == 4) op.template feed<4, true>(q, true, optr, nvalid, true);
                if (j == 5) op.template feed<5, true>(q, true, optr, nvalid, true);
                if (j == 6) op.template feed<6, true>(q, true, optr, nvalid, true);
                if (j == 7) op.template feed<7, true>(q, true, optr, nvalid, true);
#pragma unroll
                for (int k = 0; k < Op::NOUT; ++k)
                  if (Op::NOUT == 1 || optr[k]) optr[k] += p.out[k].step;
              }
            }
            continue;
          }
        }
#pragma unroll 1
        for (int g = g0; g < R / U; ++g) {
          const uint32_t rowaddr = tile + (uint32_t)(g * U) * kTileBytes + lane * kLaneBytes;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const uint4 q = lds128(rowaddr + j * kTileBytes);
            if constexpr (!OpHoistWarm<Op>::value) {
              if (c == 0 && g * U + j < 2 * HV) {
                if (j == 0) op.template warm<0>(q);
                if (j == 1) op.template warm<1>(q);
                if (j == 2) op.template warm<2>(q);
                if (j == 3) op.template warm<3>(q);
                if (j == 4) op.template warm<4>(q);
                if (j == 5) op.template warm<5>(q);
                if (j == 6) op.template warm<6>(q);
                if (j == 7) op.template warm<7>(q);
                continue;
              }
            }
            if (j == 0) op.template feed<0, true>(q, true, optr, nvalid, true);
            if (j == 1) op.template feed<1, true>(q, true, optr, nvalid, true);
            if (j == 2) op.template feed<2, true>(q, true, optr, nvalid, true);
            if (j == 3) op.template feed<3, true>(q, true, optr, nvalid, true);
            if (j == 4) op.template feed<4, true>(q, true, optr, nvalid, true);
            if (j == 5) op.template feed<5, true>(q, true, optr, nvalid, true);
            if (j == 6) op.template feed<6, true>(q, true, optr, nvalid, true);
            if (j == 7) op.template feed<7, true>(q, true, optr, nvalid, true);
#pragma unroll
            for (int k = 0; k < Op::NOUT; ++k)
              if (Op::NOUT == 1 || optr[k]) optr[k] += p.out[k].step;
          }
        }
        continue;
      }
      // first / last chunks of a band, ragged or unaligned strips: per-row tests
      constexpr int US = OpUnrollSlow<Op>::value;
#pragma unroll 1
      for (int g = 0; g < R / US; ++g) {
        const uint32_t rowaddr = tile + (uint32_t)(g * US) * kTileBytes + lane * kLaneBytes;
        // feed index fi produces output row y0 + fi - 2*HV
#pragma unroll
        for (int j = 0; j < US; ++j) {
          const int fi = c * R + g * US + j;
          if (fi < n_feed) {
            const uint4 q = lds128(rowaddr + j * kTileBytes);
            const bool emit = fi >= 2 * HV;
            if (j == 0) op.template feed<0, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 1) op.template feed<1, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 2) op.template feed<2, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 3) op.template feed<3, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 4) op.template feed<4, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 5) op.template feed<5, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 6) op.template feed<6, false>(q, emit, optr, nvalid, p.vec_store != 0);
            if (j == 7) op.template feed<7, false>(q, emit, optr, nvalid, p.vec_store != 0);
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

static inline bool aligned16(const DBatch &b) {
  return b.v.data && (((uintptr_t)b.v.data | b.v.step | b.frame_stride) & 15) == 0;
}

static inline bool strip_path_ok(const DBatch &src, int min_rows, int min_cols) {
  return aligned16(src) && src.v.rows >= min_rows && src.v.cols >= min_cols;
}

// Picks the band height.  Measured on B200 (profiles/README.md, sweep r1d): short bands win
// although every band re-feeds 2*hv warm-up rows -- the work-claim scheduler balances better
// with many small items, and the rows in flight across all resident warps then span a few
// tens of MB (TLB reach, L2-resident halo rows) instead of hundreds.  36 rows (40 fed rows =
// 5 chunks of R) measured best for the 5x5 Gaussian: 8.04 us per 4K frame vs 8.64 at 60 rows
// and 10.9 at 244.  Tiny jobs use shorter bands to occupy more warps.
static inline int pick_band_rows(Ctx *c, const char *optname, int rows, int strips, int n, int hv, int nw = kNW,
                                 int dflt = 0) {
  int64_t forced = opt_get(optname, 0);
  if (forced > 0) return (int)forced;
  const long long warps = (long long)ctx_sm_count(c) * nw;
  int br = dflt > 0 ? dflt : 5 * kR - 2 * hv;
  const long long items = (long long)strips * n * ((rows + br - 1) / br);
  if (items < warps) {
    // A job that fits one round (a single frame, a band of the host pipeline): its duration is the longest
    // item, so cut the rows into as many bands as there are warps for this many strips -- one round, every warp
    // busy, the shortest items that still are one round.  Measured on one 4K BGR frame: 14.6 us at 23 rows
    // (2256 items for 2368 warps) vs 16.7 at 28 and 20.1 at 20 (two rounds).
    const long long per = warps / ((long long)strips * n);  // bands available per strip column
    long long b = per > 0 ? (rows + per - 1) / per : br;
    if (b < 8) b = 8;
    if (b < br) br = (int)b;
  }
  return br;
}

template <class Op, int S = kS, int NW = kNW, int R = kR>
static int launch_strip(Ctx *c, const DBatch &src, const DBatch *outs, int nout, const char *band_opt,
                        cudaStream_t s, const int32_t *taps_x = nullptr, const int32_t *taps_y = nullptr,
                        const float *ftaps = nullptr, int nftaps = 0, bool wide_taps = false) {
  const CUtensorMap *tmap = nullptr;  // cached per (base, geometry): the reference API calls once per frame on reused buffers
  RCV_TRY(ctx_tmap_rows_u32(c, &tmap, src.v.data, src.v.row_bytes(), src.v.rows, src.v.step, src.n, src.frame_stride,
                            kTileBytes / 4, R));
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
  p.row_begin = 0;
  p.row_end = p.rows;
  if (outs[0].windowed()) {
    p.row_begin = outs[0].y0 < 0 ? 0 : outs[0].y0;
    p.row_end = outs[0].y1 > p.rows ? p.rows : outs[0].y1;
    if (p.row_end <= p.row_begin) return RCV_OK;
  }
  p.row_bytes = (int)src.v.row_bytes();
  p.strips = ceil_div(p.row_bytes, kTileBytes - 2 * OpHaloLanes<Op>::value * kLaneBytes);
  p.band_rows = pick_band_rows(c, band_opt, p.row_end - p.row_begin, p.strips, src.n, Op::HV, NW, OpBandRows<Op>::value);
  p.bands = ceil_div(p.row_end - p.row_begin, p.band_rows);
  p.n_frames = src.n;
  p.total_items = (long long)p.strips * p.bands * src.n;
  if (p.total_items >= (1LL << 31)) return RCV_ERR_UNSUPPORTED;  // the kernel decodes items in 32 bits
  strip_fast_div((uint32_t)p.strips, (uint64_t)p.total_items, &p.div_strips_mul, &p.div_strips_sh);
  strip_fast_div((uint32_t)p.bands, (uint64_t)p.total_items, &p.div_bands_mul, &p.div_bands_sh);
  p.next_item = nullptr;
  p.reset_item = nullptr;
  for (int i = 0; i < 4; ++i) {
    p.taps_x[i] = taps_x ? (uint32_t)taps_x[i] : 0;
    p.taps_y[i] = taps_y ? (uint32_t)taps_y[i] : 0;
  }
  for (int i = 0; i < 8; ++i) {  // wide_taps: the caller's arrays hold 8 entries each
    p.wtaps[i] = (wide_taps && taps_x) ? (uint32_t)taps_x[i] : 0;
    p.wtaps[8 + i] = (wide_taps && taps_y) ? (uint32_t)taps_y[i] : 0;
  }
  for (int i = 0; i < 52; ++i) p.ftaps[i] = (ftaps && i < nftaps) ? ftaps[i] : 0.0f;

  auto kern = k_strip<Op, R, S, NW>;
  const int smem = NW * S * R * kTileBytes + NW * S * 8;
  static bool attr_done[16] = {};  // per template instantiation, per device
  const int dev = ctx_device(c) & 15;
  if (!attr_done[dev]) {
    RCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[dev] = true;
  }
  long long blocks = (p.total_items + NW - 1) / NW;
  int64_t grid_opt = opt_get("strip.grid", 0);
  int grid = (int)(blocks < ctx_sm_count(c) ? blocks : ctx_sm_count(c));
  if (grid_opt > 0) grid = (int)grid_opt;
  if (opt_get("strip.dynamic", 1) != 0 && p.total_items > (long long)grid * NW) {
    void *ctr = nullptr;
    const bool fresh = c->scratch_bytes[SCR_COUNTER] == 0;
    RCV_TRY(ctx_scratch(c, SCR_COUNTER, 2 * sizeof(unsigned long long), &ctr));
    if (fresh) RCV_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), s));
    const unsigned which = c->counter_parity++ & 1u;
    p.next_item = (unsigned long long *)ctr + which;
    p.reset_item = (unsigned long long *)ctr + (which ^ 1u);
  }
  if (opt_get("strip.pdl", 1) != 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RCV_CUDA(cudaLaunchKernelEx(&cfg, kern, *tmap, p));
  } else {
    kern<<<grid, NW * 32, smem, s>>>(*tmap, p);
  }
  count_launch();
  RCV_CUDA(cudaGetLastError());
  return RCV_OK;
}

}  // namespace rcv
