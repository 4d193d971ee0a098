// abi.cu -- the extern "C" entry points declared in include/rcv_imgproc.h.
//
// Everything here is host plumbing: argument validation (the reference's loops
// silently return on short buffers, rustcv/src/videoio/mod.rs:346-348; this ABI returns
// a code instead), staging of host-resident Mats through device scratch, batching, and
// the synchronous-by-default contract of the reference API (README.md:31).
#include "rcv_internal.cuh"

#include <nvtx3/nvToolsExt.h>
#include <sched.h>

#include <cstring>
#include <vector>

using namespace rcv;

namespace {

// NVTX range around every host-side phase of a call (visible in Nsight Systems; a no-op without a tool attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

size_t elem_size(int depth) { return depth == RCV_F32 ? 4 : 1; }
size_t mat_row_bytes(const RcvMat *m) { return (size_t)m->cols * m->channels * elem_size(m->depth); }
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
bool is_host(const RcvMat *m) { return m->loc != RCV_DEVICE; }

int check_mat(const RcvMat *m, const char *name) {
  if (!m) return fail(RCV_ERR_ARG, "%s is NULL", name);
  if (m->rows < 0 || m->cols < 0) return fail(RCV_ERR_ARG, "%s has negative dimensions", name);
  if (m->channels < 1 || m->channels > 4) return fail(RCV_ERR_DEPTH, "%s.channels = %d (want 1..4)", name, m->channels);
  if (m->depth != RCV_U8 && m->depth != RCV_F32) return fail(RCV_ERR_DEPTH, "%s.depth = %d", name, m->depth);
  if (m->loc > RCV_HOST_PINNED) return fail(RCV_ERR_ARG, "%s.loc = %d", name, m->loc);
  if (m->rows > 0 && m->cols > 0) {
    if (!m->data) return fail(RCV_ERR_ARG, "%s.data is NULL", name);
    if (m->step < mat_row_bytes(m))
      return fail(RCV_ERR_SIZE, "%s.step %zu < row bytes %zu", name, m->step, mat_row_bytes(m));
    if (m->depth == RCV_F32 && ((m->step & 3) || ((uintptr_t)m->data & 3)))
      return fail(RCV_ERR_SIZE, "%s: f32 data/step must be 4-byte aligned", name);
  }
  return RCV_OK;
}

int check_same_size(const RcvMat *a, const RcvMat *b, const char *what) {
  if (a->rows != b->rows || a->cols != b->cols)
    return fail(RCV_ERR_SIZE, "%s: dst is %dx%d, expected %dx%d (the caller sizes dst)", what, b->rows, b->cols, a->rows,
                a->cols);
  return RCV_OK;
}

DView view_of(const RcvMat *m, void *data, size_t step) {
  DView v;
  v.data = (uint8_t *)data;
  v.rows = m->rows;
  v.cols = m->cols;
  v.step = step;
  v.cn = m->channels;
  v.depth = m->depth;
  return v;
}

DBatch single(const DView &v) {
  DBatch b;
  b.v = v;
  b.frame_stride = 0;
  b.n = 1;
  return b;
}

DBatch none_batch() {
  DBatch b;
  memset(&b, 0, sizeof(b));
  b.n = 1;
  return b;
}

// the context an op runs on: the device of any device-resident Mat, else the default
Ctx *pick_ctx(const RcvMat *const *mats, int n) {
  int dev = -1;
  for (int i = 0; i < n; ++i) {
    if (!mats[i] || mats[i]->loc != RCV_DEVICE) continue;
    if (dev >= 0 && mats[i]->device != dev) {
      set_error("Mats live on different GPUs (%d and %d)", dev, mats[i]->device);
      return nullptr;
    }
    dev = mats[i]->device;
  }
  return dev >= 0 ? ctx_get(dev) : ctx_default();
}

// bytes spanned by a Mat's rows
size_t mat_span(const RcvMat *m) {
  if (m->rows <= 0 || m->cols <= 0) return 0;
  return (size_t)(m->rows - 1) * m->step + mat_row_bytes(m);
}

// Can the copy engines DMA this host Mat directly?  Pinned storage of the library, a range the caller registered
// (rcv_host_register), or -- "host.auto_register" -- a buffer registered on first sight.  Anything else is
// pageable memory and goes through the bounce ring (pipeline_frame).
bool host_dma_ok(const RcvMat *m) {
  if (m->loc == RCV_HOST_PINNED) return true;
  const size_t span = mat_span(m);
  if (span == 0) return true;
  if (host_range_pinned(m->data, span)) return true;
  if (opt_get("host.auto_register", 0) != 0) return host_auto_register(m->data, span);
  return false;
}

// A Mat made usable by a kernel: device Mats are used in place, host Mats get device
// scratch with a 256-byte aligned pitch.
struct Staged {
  DView v;
  bool staged = false;
};

int stage_alloc(Ctx *c, const RcvMat *m, int slot, Staged *out) {
  if (!is_host(m)) {
    out->v = view_of(m, m->data, m->step);
    out->staged = false;
    return RCV_OK;
  }
  size_t pitch = align_up(mat_row_bytes(m), 256);
  if (pitch == 0) pitch = 256;
  void *p = nullptr;
  RCV_TRY(ctx_scratch(c, slot, pitch * (size_t)(m->rows > 0 ? m->rows : 1), &p));
  out->v = view_of(m, p, pitch);
  out->staged = true;
  return RCV_OK;
}

int copy_in(const RcvMat *m, const Staged &st, cudaStream_t s) {
  if (!st.staged || m->rows == 0 || m->cols == 0) return RCV_OK;
  RCV_CUDA(cudaMemcpy2DAsync(st.v.data, st.v.step, m->data, m->step, mat_row_bytes(m), m->rows,
                             cudaMemcpyHostToDevice, s));
  return RCV_OK;
}

// written_cols < cols: the op leaves the trailing columns of dst untouched (the odd last
// pixel of a YUYV row, videoio/mod.rs:350), so they must not be overwritten by scratch.
int copy_out(const RcvMat *m, const Staged &st, cudaStream_t s, int written_cols = -1) {
  if (!st.staged || m->rows == 0 || m->cols == 0) return RCV_OK;
  size_t rb = mat_row_bytes(m);
  if (written_cols >= 0 && written_cols < m->cols) rb = (size_t)written_cols * m->channels * elem_size(m->depth);
  if (rb == 0) return RCV_OK;
  RCV_CUDA(cudaMemcpy2DAsync(m->data, m->step, st.v.data, st.v.step, rb, m->rows, cudaMemcpyDeviceToHost, s));
  return RCV_OK;
}

// Zero-copy: a pinned host Mat (rcv_pinned_alloc: cudaHostAlloc, one address for CPU and GPU under UVA) can be
// read by TMA and written by the kernels' 128-bit stores directly over PCIe -- no staging copy, no HBM round
// trip, reads and writes in flight together.  Only for 16-byte aligned Mats (the TMA / vector paths); mode is
// the "host.zero_copy" option: 0 = never (stage through HBM), 1 = whenever both sides qualify.
bool zero_copy_mat(const RcvMat *m) {
  if (m->loc == RCV_DEVICE) return true;
  return m->loc == RCV_HOST_PINNED && m->rows > 0 && m->cols > 0 && ((((uintptr_t)m->data) | m->step) & 15) == 0;
}
bool zero_copy_pair(const RcvMat *src, const RcvMat *dst) {
  return opt_get("host.zero_copy", 0) != 0 && zero_copy_mat(src) && zero_copy_mat(dst);
}

// ---- the host pipeline ------------------------------------------------------------------------
// How an op may be cut into row bands, so that the H2D copy, the kernel and the D2H copy of ONE frame overlap
// (a single 4K BGR frame is 0.5 ms of PCIe time each way: run back to back the synchronous call costs the
// sum, banded it costs about the larger of the two).
struct Banding {
  int halo = -1;         // source rows needed above / below an output band; < 0: the op cannot be banded
  bool subview = false;  // pointwise op: a band is just a sub-image (no row-window support needed in the kernel)
};
constexpr Banding kNoBands = {-1, false};
inline Banding band_window(int halo) { return Banding{halo, false}; }
constexpr Banding kBandPointwise = {0, true};

int copy_in_rows(const RcvMat *m, const Staged &st, int r0, int r1, cudaStream_t s) {
  if (!st.staged || r1 <= r0 || m->cols == 0) return RCV_OK;
  if (m->step == st.v.step && m->step == mat_row_bytes(m)) {  // both sides packed at one pitch: a flat copy
    RCV_CUDA(cudaMemcpyAsync(st.v.data + (size_t)r0 * st.v.step, (const uint8_t *)m->data + (size_t)r0 * m->step,
                             (size_t)(r1 - r0) * m->step, cudaMemcpyHostToDevice, s));
    return RCV_OK;
  }
  RCV_CUDA(cudaMemcpy2DAsync(st.v.data + (size_t)r0 * st.v.step, st.v.step, (const uint8_t *)m->data + (size_t)r0 * m->step,
                             m->step, mat_row_bytes(m), r1 - r0, cudaMemcpyHostToDevice, s));
  return RCV_OK;
}

size_t written_row_bytes(const RcvMat *m, int written_cols) {
  size_t rb = mat_row_bytes(m);
  if (written_cols >= 0 && written_cols < m->cols) rb = (size_t)written_cols * m->channels * elem_size(m->depth);
  return rb;
}

int copy_out_rows(const RcvMat *m, const Staged &st, int r0, int r1, cudaStream_t s, int written_cols) {
  if (!st.staged || r1 <= r0 || m->cols == 0) return RCV_OK;
  const size_t rb = written_row_bytes(m, written_cols);
  if (rb == 0) return RCV_OK;
  if (m->step == st.v.step && m->step == rb) {
    RCV_CUDA(cudaMemcpyAsync((uint8_t *)m->data + (size_t)r0 * m->step, st.v.data + (size_t)r0 * st.v.step,
                             (size_t)(r1 - r0) * m->step, cudaMemcpyDeviceToHost, s));
    return RCV_OK;
  }
  RCV_CUDA(cudaMemcpy2DAsync((uint8_t *)m->data + (size_t)r0 * m->step, m->step, st.v.data + (size_t)r0 * st.v.step,
                             st.v.step, rb, r1 - r0, cudaMemcpyDeviceToHost, s));
  return RCV_OK;
}

// Rows per band for this src/dst pair, 0 = do not band.  Measured on a 4K BGR frame, pinned Mats
// (profiles/r1_host_pipeline.txt): every band costs ~20-40 us of stream hand-offs, so few large bands win -- 6 MB
// bands (4 per frame) 0.72-0.78 ms against 0.93 ms unbanded and 0.95 ms with 1 MB bands.  A pageable side adds two
// CPU-copy stages to the pipeline (Mat -> bounce buffer -> ... -> bounce buffer -> Mat), whose fill and drain are
// one band each: those frames use smaller bands ("host.bounce_band_bytes").
int pick_host_band_rows(const RcvMat *src, const RcvMat *dst, const Banding &bd, bool bounce) {
  if (bd.halo < 0 || src->rows != dst->rows || src->rows <= 0) return 0;
  if (!is_host(src) && !is_host(dst)) return 0;
  const int64_t band_bytes = bounce ? opt_get("host.bounce_band_bytes", 4 << 20) : opt_get("host.band_bytes", 6 << 20);
  if (band_bytes <= 0) return 0;
  size_t rb = mat_row_bytes(src) > mat_row_bytes(dst) ? mat_row_bytes(src) : mat_row_bytes(dst);
  if (rb == 0) return 0;
  int64_t br = band_bytes / (int64_t)rb;
  const int64_t min_br = (src->rows + kMaxBands - 1) / kMaxBands;  // one event per band: at most kMaxBands of them
  if (br < min_br) br = min_br;
  if (br < 8) br = 8;
  br = (br + 7) & ~(int64_t)7;
  if ((int64_t)src->rows < 2 * br) return 0;
  return (int)br;
}

int ensure_bounce(Ctx *c, void **buf, size_t *have, size_t want) {
  if (*have >= want) return RCV_OK;
  if (*buf) {
    RCV_CUDA(cudaStreamSynchronize(c->s_in));
    RCV_CUDA(cudaStreamSynchronize(c->s_out));
    RCV_CUDA(cudaFreeHost(*buf));
    *buf = nullptr;
    *have = 0;
  }
  want += want / 8 + 4096;
  RCV_CUDA(cudaHostAlloc(buf, want, cudaHostAllocPortable));
  *have = want;
  return RCV_OK;
}

void wait_drained(Ctx *c, int k) {
  while (c->out_pending[k].load(std::memory_order_acquire) > 0) sched_yield();
}

// One frame through ring slot k: H2D on s_in, kernel on the library stream, D2H on s_out, chained by events.
// band_rows > 0: the three stages run band by band (copy rows [.., r1 + halo) in, produce rows [r0, r1), copy
// them out); 0: whole frame at once.  `in` / `out` are whole-frame views (staged scratch or the Mat itself).
// src_bounce / dst_bounce: that side is pageable memory the copy engines cannot take at speed -- its rows pass
// through the slot's pinned bounce buffer: the memcpy pool fills bounce_in[k] band by band ahead of the H2D copy;
// the GPU's drain thread empties bounce_out[k] band by band behind the D2H copy (ev_band[k][b]).
template <class F>
int pipeline_frame(Ctx *c, const RcvMat *src, RcvMat *dst, const Staged &in, const Staged &out, int k, bool wait_slot,
                   int band_rows, const Banding &bd, F &launch, int written_cols, bool src_bounce, bool dst_bounce) {
  const int rows = dst->rows;
  const int br = band_rows > 0 ? band_rows : (rows > 0 ? rows : 1);
  RcvMat bsrc = *src, bdst = *dst;  // the Mats as the copy engines see them
  if (src_bounce) {
    const size_t rb = mat_row_bytes(src);
    RCV_TRY(ensure_bounce(c, &c->bounce_in[k], &c->bounce_in_bytes[k], rb * (size_t)src->rows));
    if (wait_slot) RCV_CUDA(cudaEventSynchronize(c->ev_in[k]));  // the slot's previous frame has left the buffer
    bsrc.data = c->bounce_in[k];
    bsrc.step = rb;
  }
  if (dst_bounce) {
    const size_t rb = written_row_bytes(dst, written_cols);
    wait_drained(c, k);  // the slot's previous frame has been copied out of the buffer
    RCV_TRY(ensure_bounce(c, &c->bounce_out[k], &c->bounce_out_bytes[k], mat_row_bytes(dst) * (size_t)dst->rows));
    bdst.data = c->bounce_out[k];
    bdst.step = rb ? rb : 1;
  }
  int uploaded = 0, b = 0;
  for (int r0 = 0; r0 == 0 || r0 < rows; r0 += br, ++b) {
    const int r1 = r0 + br < rows ? r0 + br : rows;
    if (r0 == 0 && wait_slot) RCV_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_out[k], 0));  // slot k fully drained
    int need = band_rows > 0 ? r1 + bd.halo : src->rows;
    if (need > src->rows) need = src->rows;
    if (need > uploaded) {
      if (src_bounce && in.staged)
        host_copy2d((uint8_t *)bsrc.data + (size_t)uploaded * bsrc.step, bsrc.step,
                    (const uint8_t *)src->data + (size_t)uploaded * src->step, src->step, mat_row_bytes(src), need - uploaded);
      RCV_TRY(copy_in_rows(&bsrc, in, uploaded, need, c->s_in));
      uploaded = need;
    }
    RCV_CUDA(cudaEventRecord(c->ev_in[k], c->s_in));
    RCV_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
    DBatch sb = single(in.v), db = single(out.v);
    if (band_rows > 0) {
      if (bd.subview) {
        sb.v.data += (size_t)r0 * sb.v.step;
        db.v.data += (size_t)r0 * db.v.step;
        sb.v.rows = db.v.rows = r1 - r0;
      } else {
        db.y0 = r0;
        db.y1 = r1;
      }
    }
    RCV_TRY(launch(c, sb, db, c->stream));
    RCV_CUDA(cudaEventRecord(c->ev_k[k], c->stream));
    RCV_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[k], 0));
    RCV_TRY(copy_out_rows(&bdst, out, r0, r1, c->s_out, written_cols));
    if (dst_bounce && out.staged && r1 > r0) {
      const size_t rb = written_row_bytes(dst, written_cols);
      if (rb) {
        RCV_CUDA(cudaEventRecord(c->ev_band[k][b % kMaxBands], c->s_out));
        c->out_pending[k].fetch_add(1, std::memory_order_acq_rel);
        drain_submit(c->device, DrainJob{c->ev_band[k][b % kMaxBands], (uint8_t *)dst->data + (size_t)r0 * dst->step, dst->step,
                                         (const uint8_t *)bdst.data + (size_t)r0 * bdst.step, bdst.step, rb, r1 - r0,
                                         &c->out_pending[k], &c->drain_failed});
      }
    }
  }
  RCV_CUDA(cudaEventRecord(c->ev_out[k], c->s_out));
  return RCV_OK;
}

int sync_pipeline(Ctx *c) {
  cudaError_t e1 = cudaStreamSynchronize(c->s_out);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaError_t e3 = cudaStreamSynchronize(c->s_in);
  for (int k = 0; k < kRing; ++k) wait_drained(c, k);  // the drain thread never blocks for long: its events have fired
  RCV_CUDA(e1);
  RCV_CUDA(e2);
  RCV_CUDA(e3);
  if (c->drain_failed.exchange(0)) return fail(RCV_ERR_CUDA, "a D2H band of the bounce pipeline failed");
  return RCV_OK;
}

// srcs[i] -> dsts[i] with at least one side in host memory: software pipeline over the staging ring.
template <class F>
int run_host_pipeline(Ctx *c, const RcvMat *srcs, RcvMat *dsts, int n, F &launch, int written_cols, const Banding &bd) {
  NvtxRange nvtx("rcv host pipeline");
  Staged si[kRing], so[kRing];
  const int depth = n < kRing ? n : kRing;
  // Direct write (option "host.direct_write", OFF by default): a pinned, 16-byte aligned destination written by
  // the kernel itself over PCIe.  Measured (profiles/r1_host_pipeline.txt): the Gaussian strip kernel alone
  // stores to host memory at 49 GB/s (whole 480-byte row segments), the rate of the D2H copy it would replace,
  // but next to a concurrent H2D copy the pair reaches only 41 GB/s each way against 44-45 GB/s for two copy
  // engines, and kernels with narrower stores (YUYV->BGR: 48 B per thread) drop to a third.  TMA reads straight
  // from host memory ("host.zero_copy") reach 34-38 GB/s.  The copy engines stay the default on both sides.
  bool direct = opt_get("host.direct_write", 0) != 0 && is_host(&dsts[0]);
  for (int i = 0; i < n && direct; ++i) direct = zero_copy_mat(&dsts[i]);
  for (int k = 0; k < depth; ++k) {
    RCV_TRY(stage_alloc(c, &srcs[k], SCR_STAGE_IN0 + k, &si[k]));
    if (direct)
      so[k].staged = false;
    else
      RCV_TRY(stage_alloc(c, &dsts[k], SCR_STAGE_OUT0 + 3 * k, &so[k]));
  }
  RCV_CUDA(cudaStreamSynchronize(c->stream));
  // pageable sides (frame 0 decides the band height; every frame is tested for itself)
  const bool bounce0 = (is_host(&srcs[0]) && !host_dma_ok(&srcs[0])) || (is_host(&dsts[0]) && !direct && !host_dma_ok(&dsts[0]));
  int band_rows = pick_host_band_rows(&srcs[0], &dsts[0], bd, bounce0);
  for (int i = 0; i < n; ++i) {
    const int k = i % kRing;
    Staged in = si[k], out = so[k];
    if (!in.staged) in.v = view_of(&srcs[i], srcs[i].data, srcs[i].step);
    if (!out.staged) out.v = view_of(&dsts[i], dsts[i].data, dsts[i].step);
    const bool sbounce = in.staged && !host_dma_ok(&srcs[i]);
    const bool dbounce = out.staged && !host_dma_ok(&dsts[i]);
    // inside a batch the frames themselves overlap and bands only add hand-offs (measured, 16 pinned 4K frames:
    // 8.85 ms unbanded, 9.51 ms with the first and last frame banded, 9.9 ms with every frame banded): bands are
    // for the single-Mat call
    const int br = n == 1 ? band_rows : 0;
    int rc = pipeline_frame(c, &srcs[i], &dsts[i], in, out, k, i >= kRing, br, bd, launch, written_cols, sbounce, dbounce);
    if (rc == RCV_ERR_UNSUPPORTED && br > 0) {
      // the op fell off its row-window capable kernel for this geometry: redo the frame unbanded
      RCV_TRY(sync_pipeline(c));
      band_rows = 0;
      rc = pipeline_frame(c, &srcs[i], &dsts[i], in, out, k, false, 0, bd, launch, written_cols, sbounce, dbounce);
    }
    if (rc != RCV_OK) {
      sync_pipeline(c);
      return rc;
    }
  }
  return sync_pipeline(c);
}

// Runs `launch(ctx, src_batch, dst_batch, stream)` for one src -> one dst.
template <class F>
int run_unary(const RcvMat *src, RcvMat *dst, F launch, int written_cols = -1, Banding bd = kNoBands) {
  const RcvMat *mats[2] = {src, dst};
  Ctx *c = pick_ctx(mats, 2);
  if (!c) return RCV_ERR_NOT_INIT;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!is_host(src) && !is_host(dst)) {
    RCV_TRY(launch(c, single(view_of(src, src->data, src->step)), single(view_of(dst, dst->data, dst->step)), c->stream));
    if (ctx_blocking()) RCV_CUDA(cudaStreamSynchronize(c->stream));
    return RCV_OK;
  }
  if (zero_copy_pair(src, dst)) {
    RCV_TRY(launch(c, single(view_of(src, src->data, src->step)), single(view_of(dst, dst->data, dst->step)), c->stream));
    RCV_CUDA(cudaStreamSynchronize(c->stream));  // host-visible results: always synchronous
    return RCV_OK;
  }
  return run_host_pipeline(c, src, dst, 1, launch, written_cols, bd);
}

// Batches.  Device Mats with a uniform frame stride: one launch.  Device Mats otherwise:
// one launch per frame.  Host Mats: H2D / kernel / D2H pipelined over a ring of kRing slots.
bool uniform_device_batch(const RcvMat *m, int n, DBatch *out) {
  for (int i = 0; i < n; ++i)
    if (m[i].loc != RCV_DEVICE || m[i].device != m[0].device) return false;
  ptrdiff_t stride = n > 1 ? (uint8_t *)m[1].data - (uint8_t *)m[0].data : 0;
  if (n > 1 && stride <= 0) return false;
  for (int i = 1; i < n; ++i)
    if ((uint8_t *)m[i].data - (uint8_t *)m[i - 1].data != stride) return false;
  out->v = view_of(&m[0], m[0].data, m[0].step);
  out->frame_stride = (size_t)stride;
  out->n = n;
  return true;
}

int check_batch_geometry(const RcvMat *m, int n, const char *name) {
  for (int i = 0; i < n; ++i) {
    RCV_TRY(check_mat(&m[i], name));
    if (m[i].rows != m[0].rows || m[i].cols != m[0].cols || m[i].channels != m[0].channels ||
        m[i].depth != m[0].depth || m[i].step != m[0].step || m[i].loc != m[0].loc)
      return fail(RCV_ERR_SIZE, "%s[%d] differs in geometry/location from %s[0]", name, i, name);
  }
  return RCV_OK;
}

// [a, a + span_a) and [b, b + span_b) share a byte (same address space: both host or both on one device)
bool mats_overlap(const RcvMat *a, const RcvMat *b) {
  if (!a || !b || (a->loc == RCV_DEVICE) != (b->loc == RCV_DEVICE)) return false;
  if (a->loc == RCV_DEVICE && a->device != b->device) return false;
  const size_t sa = mat_span(a), sb = mat_span(b);
  if (!sa || !sb) return false;
  const uintptr_t pa = (uintptr_t)a->data, pb = (uintptr_t)b->data;
  return pa < pb + sb && pb < pa + sa;
}

// No op here runs in place: a dst that shares bytes with ANY src of the call would race with the kernels'
// reads (within a frame, and across the frames of a batch, which run concurrently).
int check_no_alias(const RcvMat *srcs, const RcvMat *dsts, int n, const char *what) {
  for (int i = 0; i < n; ++i) {
    if (mats_overlap(&srcs[i], &dsts[i])) return fail(RCV_ERR_ARG, "%s: in-place operation is not supported (src[%d] and dst[%d] overlap)", what, i, i);
    if (i > 0 && mats_overlap(&dsts[i], &dsts[i - 1])) return fail(RCV_ERR_ARG, "%s: dst[%d] overlaps dst[%d]", what, i, i - 1);
  }
  if (n > 1 && n <= 64) {  // the quadratic check, for batches where it is free
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        if (i != j && mats_overlap(&srcs[i], &dsts[j])) return fail(RCV_ERR_ARG, "%s: dst[%d] overlaps src[%d]", what, j, i);
  }
  return RCV_OK;
}

template <class F>
int run_batch_on(Ctx *c, const RcvMat *srcs, RcvMat *dsts, int n, F &launch, int written_cols, const Banding &bd) {
  std::lock_guard<std::mutex> lk(c->mu);
  const bool src_host = is_host(&srcs[0]), dst_host = is_host(&dsts[0]);
  if (!src_host && !dst_host) {
    DBatch sb, db;
    if (uniform_device_batch(srcs, n, &sb) && uniform_device_batch(dsts, n, &db)) {
      RCV_TRY(launch(c, sb, db, c->stream));
    } else {
      for (int i = 0; i < n; ++i)
        RCV_TRY(launch(c, single(view_of(&srcs[i], srcs[i].data, srcs[i].step)),
                       single(view_of(&dsts[i], dsts[i].data, dsts[i].step)), c->stream));
    }
    if (ctx_blocking()) RCV_CUDA(cudaStreamSynchronize(c->stream));
    return RCV_OK;
  }
  if (zero_copy_pair(&srcs[0], &dsts[0])) {  // geometry and location are uniform over the batch (checked)
    bool all = true;
    for (int i = 0; i < n; ++i) all = all && zero_copy_mat(&srcs[i]) && zero_copy_mat(&dsts[i]);
    if (all) {
      for (int i = 0; i < n; ++i)
        RCV_TRY(launch(c, single(view_of(&srcs[i], srcs[i].data, srcs[i].step)),
                       single(view_of(&dsts[i], dsts[i].data, dsts[i].step)), c->stream));
      RCV_CUDA(cudaStreamSynchronize(c->stream));
      return RCV_OK;
    }
  }
  return run_host_pipeline(c, srcs, dsts, n, launch, written_cols, bd);
}

// One GPU's share of a multi-GPU batch, run on that GPU's worker thread (multi.cu).
template <class F>
struct MultiPart {
  Ctx *c;
  std::vector<RcvMat> s, d;
  F *launch;
  int written_cols;
  Banding bd;
  static int run(void *p) {
    MultiPart *m = (MultiPart *)p;
    NvtxRange nvtx("rcv multi-GPU share");
    cudaSetDevice(m->c->device);
    return run_batch_on(m->c, m->s.data(), m->d.data(), (int)m->s.size(), *m->launch, m->written_cols, m->bd);
  }
};

// ngpus == 0: the classic single-GPU batch (the GPU of the device Mats, else the default GPU).
// ngpus != 0 (the *_batch_multi entry points; < 0 = every initialised GPU): frames are independent, so host
// frames go j -> GPU j mod N and device frames to the GPU that owns them; each GPU's share runs on its own worker
// thread with its own streams and staging ring, and the calling thread waits for all of them.
template <class F>
int run_batch(const RcvMat *srcs, RcvMat *dsts, int n, F launch, int written_cols = -1, Banding bd = kNoBands, int ngpus = 0) {
  if (n == 0) return RCV_OK;
  if (ngpus == 0) {
    std::vector<const RcvMat *> mats;
    for (int i = 0; i < n; ++i) {
      mats.push_back(&srcs[i]);
      mats.push_back(&dsts[i]);
    }
    Ctx *c = pick_ctx(mats.data(), (int)mats.size());
    if (!c) return RCV_ERR_NOT_INIT;
    return run_batch_on(c, srcs, dsts, n, launch, written_cols, bd);
  }
  NvtxRange nvtx("rcv multi-GPU batch");
  int devs[16];
  int have = devices_initialised(devs, 16);
  if (have == 0) return fail(RCV_ERR_NOT_INIT, "rcv_init has not been called (no CPU fallback exists)");
  if (ngpus < 0) ngpus = have;
  if (ngpus > have) return fail(RCV_ERR_ARG, "ngpus = %d but only %d GPU(s) are initialised (rcv_init_multi)", ngpus, have);
  std::vector<MultiPart<F>> parts(ngpus);
  for (int g = 0; g < ngpus; ++g) {
    parts[g].c = ctx_get(devs[g]);
    if (!parts[g].c) return RCV_ERR_NOT_INIT;
    parts[g].launch = &launch;
    parts[g].written_cols = written_cols;
    parts[g].bd = bd;
  }
  for (int i = 0; i < n; ++i) {
    int g = i % ngpus;
    int dev = -1;
    if (srcs[i].loc == RCV_DEVICE) dev = srcs[i].device;
    if (dsts[i].loc == RCV_DEVICE) {
      if (dev >= 0 && dsts[i].device != dev)
        return fail(RCV_ERR_ARG, "frame %d: src lives on GPU %d, dst on GPU %d (no inter-GPU traffic on the pixel path)", i, dev, dsts[i].device);
      dev = dsts[i].device;
    }
    if (dev >= 0) {
      g = -1;
      for (int q = 0; q < ngpus; ++q)
        if (devs[q] == dev) g = q;
      if (g < 0) return fail(RCV_ERR_ARG, "frame %d lives on GPU %d, which is not among the %d GPUs of this call", i, dev, ngpus);
    }
    parts[g].s.push_back(srcs[i]);
    parts[g].d.push_back(dsts[i]);
  }
  int njobs = 0;
  for (int g = 0; g < ngpus; ++g) njobs += parts[g].s.empty() ? 0 : 1;
  MultiJoin *join = multi_begin(njobs);
  for (int g = 0, slot = 0; g < ngpus; ++g)
    if (!parts[g].s.empty()) multi_submit(join, devs[g], slot++, &MultiPart<F>::run, &parts[g]);
  return multi_wait(join);
}

// argument checks shared by every batch entry point: arrays present, one geometry / location per side, and no
// dst sharing bytes with any src of the call (every element, not just element 0)
int check_batch(const RcvMat *srcs, const RcvMat *dsts, int n, const char *what) {
  if (n < 0 || (n > 0 && (!srcs || !dsts))) return fail(RCV_ERR_ARG, "bad batch arguments");
  if (n == 0) return RCV_OK;
  RCV_TRY(check_batch_geometry(srcs, n, "srcs"));
  RCV_TRY(check_batch_geometry(dsts, n, "dsts"));
  return check_no_alias(srcs, dsts, n, what);
}

int check_cvt(const RcvMat *src, const RcvMat *dst, int code) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(dst, "dst"));
  int sc = 0, dc = 0;
  switch (code) {
    case RCV_COLOR_YUYV2BGR:
    case RCV_COLOR_UYVY2BGR: sc = 2; dc = 3; break;
    case RCV_COLOR_BGRA2BGR: sc = 4; dc = 3; break;
    case RCV_COLOR_RGB2BGR: sc = 3; dc = 3; break;
    case RCV_COLOR_BGR2GRAY: sc = 3; dc = 1; break;
    case RCV_COLOR_BGR2XRGB32: sc = 3; dc = 4; break;
    case RCV_COLOR_YUYV2GRAY: sc = 2; dc = 1; break;
    default: return fail(RCV_ERR_ARG, "unknown colour conversion code %d", code);
  }
  if (src->depth != RCV_U8 || dst->depth != RCV_U8) return fail(RCV_ERR_DEPTH, "cvtColor: u8 only");
  if (src->channels != sc || dst->channels != dc)
    return fail(RCV_ERR_DEPTH, "cvtColor code %d wants %d -> %d channels, got %d -> %d", code, sc, dc, src->channels,
                dst->channels);
  if (code == RCV_COLOR_BGR2XRGB32 && dst->rows > 0 && ((dst->step & 3) || ((uintptr_t)dst->data & 3)))
    return fail(RCV_ERR_SIZE, "XRGB32 dst must be 4-byte aligned");
  // the vector kernels read and write different byte ranges per thread: overlapping buffers would race
  if (mats_overlap(src, dst)) return fail(RCV_ERR_ARG, "cvtColor: in-place operation is not supported (src and dst overlap)");
  return check_same_size(src, dst, "cvtColor");
}

// 4:2:2 formats convert cols/2 macro-pixels per row: an odd last column is not written
int cvt_written_cols(const RcvMat *src, int code) {
  if (code == RCV_COLOR_YUYV2BGR || code == RCV_COLOR_UYVY2BGR || code == RCV_COLOR_YUYV2GRAY) return src->cols & ~1;
  return -1;
}

int check_filter_pair(const RcvMat *src, const RcvMat *dst, const char *what) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(dst, "dst"));
  if (src->depth != dst->depth || src->channels != dst->channels)
    return fail(RCV_ERR_DEPTH, "%s: dst depth/channels differ from src", what);
  if (mats_overlap(src, dst)) return fail(RCV_ERR_ARG, "%s: in-place operation is not supported", what);
  return check_same_size(src, dst, what);
}

}  // namespace

extern "C" {

// ---- storage ---------------------------------------------------------------------------
int rcv_mat_alloc_device_batch(RcvMat *mats, int32_t n, int32_t rows, int32_t cols, int32_t channels, int32_t depth,
                               int32_t device) {
  if (!mats || n < 1) return fail(RCV_ERR_ARG, "mats is NULL or n < 1");
  if (rows < 0 || cols < 0 || channels < 1 || channels > 4 || (depth != RCV_U8 && depth != RCV_F32))
    return fail(RCV_ERR_ARG, "bad geometry %dx%dx%d depth %d", rows, cols, channels, depth);
  Ctx *c = device < 0 ? ctx_default() : ctx_get(device);
  if (!c) return RCV_ERR_NOT_INIT;
  size_t rb = (size_t)cols * channels * elem_size(depth);
  size_t step = align_up(rb ? rb : 1, 256);
  size_t frame = step * (size_t)(rows ? rows : 1);
  void *p = nullptr;
  RCV_CUDA(cudaMalloc(&p, frame * n));
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->dev_allocs.insert(p);
  }
  for (int i = 0; i < n; ++i) {
    mats[i].data = (uint8_t *)p + frame * i;
    mats[i].rows = rows;
    mats[i].cols = cols;
    mats[i].step = step;
    mats[i].channels = (uint8_t)channels;
    mats[i].depth = (uint8_t)depth;
    mats[i].loc = RCV_DEVICE;
    mats[i].reserved = 0;
    mats[i].device = c->device;
  }
  return RCV_OK;
}

int rcv_mat_alloc_device(RcvMat *m, int32_t rows, int32_t cols, int32_t channels, int32_t depth, int32_t device) {
  return rcv_mat_alloc_device_batch(m, 1, rows, cols, channels, depth, device);
}

int rcv_mat_free_device_batch(RcvMat *mats, int32_t n) {
  if (!mats || n < 1) return fail(RCV_ERR_ARG, "mats is NULL or n < 1");
  if (mats[0].loc != RCV_DEVICE) return fail(RCV_ERR_ARG, "not a device Mat");
  Ctx *c = ctx_get(mats[0].device);
  if (!c) return RCV_ERR_NOT_INIT;
  if (mats[0].data) {
    {
      // only what rcv_mat_alloc_device[_batch] returned may be freed: a Mat carved from a batch (index > 0)
      // points into the middle of an allocation
      std::lock_guard<std::mutex> lk(c->mu);
      auto it = c->dev_allocs.find(mats[0].data);
      if (it == c->dev_allocs.end())
        return fail(RCV_ERR_ARG, "%p is not the base of a device allocation of this library (a Mat carved from a batch is "
                                 "freed with its batch)", mats[0].data);
      c->dev_allocs.erase(it);
    }
    RCV_CUDA(cudaStreamSynchronize(c->stream));
    RCV_CUDA(cudaFree(mats[0].data));
  }
  for (int i = 0; i < n; ++i) mats[i].data = nullptr;
  return RCV_OK;
}

int rcv_mat_free_device(RcvMat *m) { return rcv_mat_free_device_batch(m, 1); }

static int copy_mat(const RcvMat *src, RcvMat *dst, cudaMemcpyKind kind, int device) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(dst, "dst"));
  if (src->rows != dst->rows || src->cols != dst->cols || src->channels != dst->channels || src->depth != dst->depth)
    return fail(RCV_ERR_SIZE, "upload/download: geometry differs");
  Ctx *c = ctx_get(device);
  if (!c) return RCV_ERR_NOT_INIT;
  if (src->rows == 0 || src->cols == 0) return RCV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  RCV_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, src->data, src->step, mat_row_bytes(src), src->rows, kind, c->stream));
  RCV_CUDA(cudaStreamSynchronize(c->stream));
  return RCV_OK;
}

int rcv_mat_upload(const RcvMat *host, RcvMat *dev) {
  if (!host || !dev) return fail(RCV_ERR_ARG, "NULL Mat");
  if (host->loc == RCV_DEVICE || dev->loc != RCV_DEVICE) return fail(RCV_ERR_ARG, "upload wants host -> device");
  return copy_mat(host, dev, cudaMemcpyHostToDevice, dev->device);
}

int rcv_mat_download(const RcvMat *dev, RcvMat *host) {
  if (!host || !dev) return fail(RCV_ERR_ARG, "NULL Mat");
  if (host->loc == RCV_DEVICE || dev->loc != RCV_DEVICE) return fail(RCV_ERR_ARG, "download wants device -> host");
  return copy_mat(dev, host, cudaMemcpyDeviceToHost, dev->device);
}

// ---- conversions -------------------------------------------------------------------------
int rcv_cvt_color(const RcvMat *src, RcvMat *dst, int32_t code) {
  RCV_TRY(check_cvt(src, dst, code));
  return run_unary(src, dst, [code](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_cvt(c, s, d, code, st);
  }, cvt_written_cols(src, code), kBandPointwise);
}

static int cvt_color_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t code, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "cvtColor"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_cvt(&srcs[0], &dsts[0], code));
  return run_batch(srcs, dsts, n, [code](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_cvt(c, s, d, code, st);
  }, cvt_written_cols(&srcs[0], code), kBandPointwise, ngpus);
}
int rcv_cvt_color_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t code) {
  return cvt_color_batch(srcs, dsts, n, code, 0);
}
int rcv_cvt_color_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, int32_t code) {
  return cvt_color_batch(srcs, dsts, n, code, ngpus > 0 ? ngpus : -1);
}

int rcv_yuyv_to_bgr(const RcvMat *src, RcvMat *dst) { return rcv_cvt_color(src, dst, RCV_COLOR_YUYV2BGR); }

static int packed_cvt(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len, size_t width, size_t height,
                      int code) {
  if (!src || !dst) return fail(RCV_ERR_ARG, "NULL buffer");
  if (width > 0x7fffffffu / 4 || height > 0x7fffffffu || width * height > 0x7fffffffu / 4)
    return fail(RCV_ERR_SIZE, "frame too large");
  const size_t px = width * height;
  RcvMat s, d;
  memset(&s, 0, sizeof(s));
  memset(&d, 0, sizeof(d));
  if (code == RCV_COLOR_YUYV2BGR) {
    // videoio/mod.rs:345-349: needs width*height*2 source bytes; converts width*height/2
    // macro-pixels as ONE packed run (stride ignored), writing 6 bytes each.
    const size_t pairs = px / 2;
    if (src_len < px * 2) return fail(RCV_ERR_SIZE, "src has %zu bytes, frame needs %zu", src_len, px * 2);
    if (dst_len < pairs * 6) return fail(RCV_ERR_SIZE, "dst has %zu bytes, needs %zu", dst_len, pairs * 6);
    s.rows = 1;
    s.cols = (int32_t)(pairs * 2);
    s.channels = 2;
    d.channels = 3;
  } else {
    if (src_len < px * 4 || dst_len < px * 3) return fail(RCV_ERR_SIZE, "buffer shorter than the %zux%zu frame", width, height);
    s.rows = 1;
    s.cols = (int32_t)px;
    s.channels = 4;
    d.channels = 3;
  }
  s.data = const_cast<uint8_t *>(src);
  s.step = (size_t)s.cols * s.channels;
  s.loc = RCV_HOST;
  d.data = dst;
  d.rows = 1;
  d.cols = s.cols;
  d.step = (size_t)d.cols * 3;
  d.loc = RCV_HOST;
  if (s.cols == 0) return RCV_OK;
  return rcv_cvt_color(&s, &d, code);
}

int rcv_yuyv_to_bgr_packed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len, size_t width,
                           size_t height) {
  return packed_cvt(src, src_len, dst, dst_len, width, height, RCV_COLOR_YUYV2BGR);
}

int rcv_bgra_to_bgr_packed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len, size_t width,
                           size_t height) {
  return packed_cvt(src, src_len, dst, dst_len, width, height, RCV_COLOR_BGRA2BGR);
}

int rcv_nv12_to_bgr(const RcvMat *y, const RcvMat *uv, RcvMat *dst) {
  RCV_TRY(check_mat(y, "y"));
  RCV_TRY(check_mat(uv, "uv"));
  RCV_TRY(check_mat(dst, "dst"));
  if (y->depth != RCV_U8 || uv->depth != RCV_U8 || dst->depth != RCV_U8) return fail(RCV_ERR_DEPTH, "NV12: u8 only");
  if (y->channels != 1 || uv->channels != 2 || dst->channels != 3) return fail(RCV_ERR_DEPTH, "NV12: want y c1, uv c2, dst c3");
  if ((y->cols & 1) || (y->rows & 1)) return fail(RCV_ERR_SIZE, "NV12 needs even dimensions");
  if (uv->rows != y->rows / 2 || uv->cols != y->cols / 2) return fail(RCV_ERR_SIZE, "uv plane must be rows/2 x cols/2");
  RCV_TRY(check_same_size(y, dst, "nv12"));
  const RcvMat *mats[3] = {y, uv, dst};
  Ctx *c = pick_ctx(mats, 3);
  if (!c) return RCV_ERR_NOT_INIT;
  std::lock_guard<std::mutex> lk(c->mu);
  Staged sy, suv, so;
  RCV_TRY(stage_alloc(c, y, SCR_STAGE_IN0, &sy));
  RCV_TRY(stage_alloc(c, uv, SCR_STAGE_IN0 + 1, &suv));
  RCV_TRY(stage_alloc(c, dst, SCR_STAGE_OUT0, &so));
  RCV_TRY(copy_in(y, sy, c->stream));
  RCV_TRY(copy_in(uv, suv, c->stream));
  RCV_TRY(launch_nv12(c, sy.v, suv.v, so.v, c->stream));
  RCV_TRY(copy_out(dst, so, c->stream));
  if (ctx_blocking() || sy.staged || suv.staged || so.staged) RCV_CUDA(cudaStreamSynchronize(c->stream));
  return RCV_OK;
}

int rcv_convert_to(const RcvMat *src, RcvMat *dst, double alpha, double beta) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(dst, "dst"));
  if (src->channels != dst->channels) return fail(RCV_ERR_DEPTH, "convertTo: channel counts differ");
  if (mats_overlap(src, dst) && (src->depth != dst->depth || src->data != dst->data || src->step != dst->step))
    return fail(RCV_ERR_ARG, "convertTo: only a same-depth conversion exactly in place may share its buffer");
  RCV_TRY(check_same_size(src, dst, "convertTo"));
  return run_unary(src, dst, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_convert(c, s, d, alpha, beta, st);
  }, -1, kBandPointwise);
}

// ---- filters -------------------------------------------------------------------------------
// rows of source a band of the banded host pipeline needs above / below itself: half the vertical kernel size the
// launcher will derive (cv::GaussianBlur: kh <= 0 means "from sigma")
static int gaussian_halo_rows(const RcvMat *src, int kh, double sigma_x, double sigma_y) {
  if (sigma_y <= 0) sigma_y = sigma_x;
  if (kh <= 0 && sigma_y > 0) kh = gaussian_ksize(sigma_y, src->depth == RCV_U8);
  return kh > 0 ? kh / 2 : 0;
}

int rcv_gaussian_blur(const RcvMat *src, RcvMat *dst, int32_t kw, int32_t kh, double sigma_x, double sigma_y) {
  RCV_TRY(check_filter_pair(src, dst, "GaussianBlur"));
  return run_unary(src, dst, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_gaussian(c, s, d, kw, kh, sigma_x, sigma_y, st);
  }, -1, band_window(gaussian_halo_rows(src, kh, sigma_x, sigma_y)));
}

static int gaussian_blur_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t kw, int32_t kh, double sigma_x,
                               double sigma_y, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "GaussianBlur"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_filter_pair(&srcs[0], &dsts[0], "GaussianBlur"));
  return run_batch(srcs, dsts, n, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_gaussian(c, s, d, kw, kh, sigma_x, sigma_y, st);
  }, -1, band_window(gaussian_halo_rows(&srcs[0], kh, sigma_x, sigma_y)), ngpus);
}
int rcv_gaussian_blur_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t kw, int32_t kh, double sigma_x,
                            double sigma_y) {
  return gaussian_blur_batch(srcs, dsts, n, kw, kh, sigma_x, sigma_y, 0);
}
int rcv_gaussian_blur_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, int32_t kw, int32_t kh,
                                  double sigma_x, double sigma_y) {
  return gaussian_blur_batch(srcs, dsts, n, kw, kh, sigma_x, sigma_y, ngpus > 0 ? ngpus : -1);
}

int rcv_sep_filter2d(const RcvMat *src, RcvMat *dst, const float *kx, int32_t kw, const float *ky, int32_t kh) {
  RCV_TRY(check_filter_pair(src, dst, "sepFilter2D"));
  if (!kx || !ky) return fail(RCV_ERR_ARG, "NULL taps");
  if (src->depth != RCV_F32) return fail(RCV_ERR_DEPTH, "rcv_sep_filter2d is f32; use rcv_sep_filter2d_q8 for u8");
  return run_unary(src, dst, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_sepfilter_f32(c, s, d, kx, kw, ky, kh, st);
  }, -1, band_window(kh / 2));
}

int rcv_sep_filter2d_q8(const RcvMat *src, RcvMat *dst, const int32_t *kx, int32_t kw, const int32_t *ky, int32_t kh) {
  RCV_TRY(check_filter_pair(src, dst, "sepFilter2D"));
  if (!kx || !ky) return fail(RCV_ERR_ARG, "NULL taps");
  if (src->depth != RCV_U8) return fail(RCV_ERR_DEPTH, "rcv_sep_filter2d_q8 is u8");
  return run_unary(src, dst, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_sepfilter_q8(c, s, d, kx, kw, ky, kh, st);
  }, -1, band_window(kh / 2));
}

// kx == ky == NULL: every GPU launches with the taps IT holds from rcv_set_kernel_broadcast (kw x-taps, then kh
// y-taps, integral values) -- the consumer of the path's one collective.
static int sep_filter2d_q8_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const int32_t *kx, int32_t kw,
                                 const int32_t *ky, int32_t kh, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "sepFilter2D"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_filter_pair(&srcs[0], &dsts[0], "sepFilter2D"));
  if (srcs[0].depth != RCV_U8) return fail(RCV_ERR_DEPTH, "rcv_sep_filter2d_q8 is u8");
  if ((kx == nullptr) != (ky == nullptr)) return fail(RCV_ERR_ARG, "kx and ky must both be given or both be NULL");
  if (kw < 1 || kh < 1 || kw > 31 || kh > 31) return fail(RCV_ERR_ARG, "kernel size %dx%d outside 1..31", kw, kh);
  return run_batch(srcs, dsts, n, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    if (kx) return launch_sepfilter_q8(c, s, d, kx, kw, ky, kh, st);
    if (c->n_coeffs < kw + kh)
      return fail(RCV_ERR_ARG, "GPU %d holds %d broadcast coefficients, the call needs %d (rcv_set_kernel_broadcast)", c->device,
                  c->n_coeffs, kw + kh);
    int32_t bx[32], by[32];
    for (int i = 0; i < kw; ++i) bx[i] = (int32_t)lrintf(c->coeffs[i]);
    for (int i = 0; i < kh; ++i) by[i] = (int32_t)lrintf(c->coeffs[kw + i]);
    return launch_sepfilter_q8(c, s, d, bx, kw, by, kh, st);
  }, -1, band_window(kh / 2), ngpus);
}
int rcv_sep_filter2d_q8_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const int32_t *kx, int32_t kw,
                              const int32_t *ky, int32_t kh) {
  return sep_filter2d_q8_batch(srcs, dsts, n, kx, kw, ky, kh, 0);
}
int rcv_sep_filter2d_q8_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, const int32_t *kx,
                                    int32_t kw, const int32_t *ky, int32_t kh) {
  return sep_filter2d_q8_batch(srcs, dsts, n, kx, kw, ky, kh, ngpus > 0 ? ngpus : -1);
}

int rcv_filter2d(const RcvMat *src, RcvMat *dst, const float *kernel, int32_t kw, int32_t kh, float delta) {
  RCV_TRY(check_filter_pair(src, dst, "filter2D"));
  if (!kernel) return fail(RCV_ERR_ARG, "NULL kernel");
  return run_unary(src, dst, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_filter2d(c, s, d, kernel, kw, kh, delta, st);
  }, -1, band_window(kh / 2));
}

static int filter2d_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const float *kernel, int32_t kw, int32_t kh,
                          float delta, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "filter2D"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_filter_pair(&srcs[0], &dsts[0], "filter2D"));
  if (!kernel) return fail(RCV_ERR_ARG, "NULL kernel");
  return run_batch(srcs, dsts, n, [=](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_filter2d(c, s, d, kernel, kw, kh, delta, st);
  }, -1, band_window(kh / 2), ngpus);
}
int rcv_filter2d_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const float *kernel, int32_t kw, int32_t kh, float delta) {
  return filter2d_batch(srcs, dsts, n, kernel, kw, kh, delta, 0);
}
int rcv_filter2d_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, const float *kernel, int32_t kw,
                             int32_t kh, float delta) {
  return filter2d_batch(srcs, dsts, n, kernel, kw, kh, delta, ngpus > 0 ? ngpus : -1);
}

static int check_sobel(const RcvMat *src, const RcvMat *mag, const RcvMat *gx, const RcvMat *gy) {
  RCV_TRY(check_mat(src, "src"));
  if (src->depth != RCV_F32 || src->channels != 1) return fail(RCV_ERR_DEPTH, "Sobel: single-channel f32 only");
  if (!mag && !gx && !gy) return fail(RCV_ERR_ARG, "Sobel: no output requested");
  const RcvMat *o[3] = {mag, gx, gy};
  for (int k = 0; k < 3; ++k) {
    if (!o[k]) continue;
    RCV_TRY(check_mat(o[k], "out"));
    if (o[k]->depth != RCV_F32 || o[k]->channels != 1) return fail(RCV_ERR_DEPTH, "Sobel: outputs are single-channel f32");
    RCV_TRY(check_same_size(src, o[k], "Sobel"));
    if (mats_overlap(src, o[k])) return fail(RCV_ERR_ARG, "Sobel: in-place operation is not supported");
  }
  return RCV_OK;
}

int rcv_sobel_mag(const RcvMat *src, RcvMat *mag, RcvMat *gx, RcvMat *gy) {
  RCV_TRY(check_sobel(src, mag, gx, gy));
  if (mag && !gx && !gy)  // the common case: one output, banded host pipeline
    return run_unary(src, mag, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
      return launch_sobel(c, s, d, none_batch(), none_batch(), st);
    }, -1, band_window(1));
  const RcvMat *mats[4] = {src, mag, gx, gy};
  Ctx *c = pick_ctx(mats, 4);
  if (!c) return RCV_ERR_NOT_INIT;
  std::lock_guard<std::mutex> lk(c->mu);
  Staged si, so[3];
  RcvMat *o[3] = {mag, gx, gy};
  DBatch ob[3];
  bool any_staged = false;
  RCV_TRY(stage_alloc(c, src, SCR_STAGE_IN0, &si));
  for (int k = 0; k < 3; ++k) {
    ob[k] = none_batch();
    if (!o[k]) continue;
    RCV_TRY(stage_alloc(c, o[k], SCR_STAGE_OUT0 + k, &so[k]));
    ob[k] = single(so[k].v);
    any_staged |= so[k].staged;
  }
  RCV_TRY(copy_in(src, si, c->stream));
  RCV_TRY(launch_sobel(c, single(si.v), ob[0], ob[1], ob[2], c->stream));
  for (int k = 0; k < 3; ++k)
    if (o[k]) RCV_TRY(copy_out(o[k], so[k], c->stream));
  if (ctx_blocking() || si.staged || any_staged) RCV_CUDA(cudaStreamSynchronize(c->stream));
  return RCV_OK;
}

static int sobel_mag_batch(const RcvMat *srcs, RcvMat *mags, int32_t n, int ngpus) {
  RCV_TRY(check_batch(srcs, mags, n, "Sobel"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_sobel(&srcs[0], &mags[0], nullptr, nullptr));
  return run_batch(srcs, mags, n, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_sobel(c, s, d, none_batch(), none_batch(), st);
  }, -1, band_window(1), ngpus);
}
int rcv_sobel_mag_batch(const RcvMat *srcs, RcvMat *mags, int32_t n) { return sobel_mag_batch(srcs, mags, n, 0); }
int rcv_sobel_mag_batch_multi(const RcvMat *srcs, RcvMat *mags, int32_t n, int32_t ngpus) {
  return sobel_mag_batch(srcs, mags, n, ngpus > 0 ? ngpus : -1);
}

// ---- geometry --------------------------------------------------------------------------------
static int check_resize(const RcvMat *src, const RcvMat *dst) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(dst, "dst"));
  if (src->depth != dst->depth || src->channels != dst->channels)
    return fail(RCV_ERR_DEPTH, "resize: dst depth/channels differ from src");
  if (dst->rows > 0 && dst->cols > 0 && (src->rows == 0 || src->cols == 0))
    return fail(RCV_ERR_SIZE, "resize from an empty image");
  if (mats_overlap(src, dst)) return fail(RCV_ERR_ARG, "resize: in-place operation is not supported");
  return RCV_OK;
}

int rcv_resize_bilinear(const RcvMat *src, RcvMat *dst) {
  RCV_TRY(check_resize(src, dst));
  return run_unary(src, dst, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_resize(c, s, d, st);
  });
}

static int resize_bilinear_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "resize"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_resize(&srcs[0], &dsts[0]));
  return run_batch(srcs, dsts, n, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_resize(c, s, d, st);
  }, -1, kNoBands, ngpus);
}
int rcv_resize_bilinear_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n) { return resize_bilinear_batch(srcs, dsts, n, 0); }
int rcv_resize_bilinear_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus) {
  return resize_bilinear_batch(srcs, dsts, n, ngpus > 0 ? ngpus : -1);
}

static int check_warp(const RcvMat *src, const RcvMat *dst, const double M[6], int inverse_map, double iM[6]) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(dst, "dst"));
  if (!M) return fail(RCV_ERR_ARG, "M is NULL");
  if (src->depth != dst->depth || src->channels != dst->channels)
    return fail(RCV_ERR_DEPTH, "warpAffine: dst depth/channels differ from src");
  if (src->depth == RCV_F32 && src->channels != 1) return fail(RCV_ERR_DEPTH, "warpAffine f32: single channel only");
  if (mats_overlap(src, dst)) return fail(RCV_ERR_ARG, "warpAffine: in-place operation is not supported");
  if (inverse_map) {
    for (int i = 0; i < 6; ++i) iM[i] = M[i];
  } else if (invert_affine(M, iM) != 0) {
    return fail(RCV_ERR_ARG, "warpAffine: singular matrix");
  }
  return RCV_OK;
}

int rcv_warp_affine(const RcvMat *src, RcvMat *dst, const double M[6], int32_t inverse_map, double border_value) {
  double iM[6];
  RCV_TRY(check_warp(src, dst, M, inverse_map, iM));
  return run_unary(src, dst, [&](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_warp_affine(c, s, d, iM, border_value, st);
  });
}

static int warp_affine_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const double M[6], int32_t inverse_map,
                             double border_value, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "warpAffine"));
  if (n == 0) return RCV_OK;
  double iM[6];
  RCV_TRY(check_warp(&srcs[0], &dsts[0], M, inverse_map, iM));
  return run_batch(srcs, dsts, n, [&](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_warp_affine(c, s, d, iM, border_value, st);
  }, -1, kNoBands, ngpus);
}
int rcv_warp_affine_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, const double M[6], int32_t inverse_map,
                          double border_value) {
  return warp_affine_batch(srcs, dsts, n, M, inverse_map, border_value, 0);
}
int rcv_warp_affine_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus, const double M[6],
                                int32_t inverse_map, double border_value) {
  return warp_affine_batch(srcs, dsts, n, M, inverse_map, border_value, ngpus > 0 ? ngpus : -1);
}

int rcv_get_rotation_matrix_2d(double cx, double cy, double angle_deg, double scale, double M[6]) {
  if (!M) return fail(RCV_ERR_ARG, "M is NULL");
  rotation_matrix(cx, cy, angle_deg, scale, M);
  return RCV_OK;
}

int rcv_invert_affine(const double M[6], double iM[6]) {
  if (!M || !iM) return fail(RCV_ERR_ARG, "NULL matrix");
  if (invert_affine(M, iM) != 0) return fail(RCV_ERR_ARG, "singular matrix");
  return RCV_OK;
}

// ---- fused chains ------------------------------------------------------------------------------
// An odd width would leave the BGR intermediate's last column unconverted (cols/2 macro-pixels per row,
// videoio/mod.rs:350) and the blur would then read it: rejected, as rcv_yuyv_to_sobel_mag does.
static int check_yuyv_gauss5(const RcvMat *src, const RcvMat *dst) {
  RCV_TRY(check_cvt(src, dst, RCV_COLOR_YUYV2BGR));
  if (src->cols & 1) return fail(RCV_ERR_SIZE, "YUYV width %d is odd", src->cols);
  return RCV_OK;
}

int rcv_yuyv_to_bgr_gaussian5(const RcvMat *src, RcvMat *dst) {
  RCV_TRY(check_yuyv_gauss5(src, dst));
  return run_unary(src, dst, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_yuyv_gauss5(c, s, d, st);
  }, -1, band_window(2));
}

static int yuyv_to_bgr_gaussian5_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n, int ngpus) {
  RCV_TRY(check_batch(srcs, dsts, n, "YUYV->GaussianBlur"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_yuyv_gauss5(&srcs[0], &dsts[0]));
  return run_batch(srcs, dsts, n, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_yuyv_gauss5(c, s, d, st);
  }, -1, band_window(2), ngpus);
}
int rcv_yuyv_to_bgr_gaussian5_batch(const RcvMat *srcs, RcvMat *dsts, int32_t n) {
  return yuyv_to_bgr_gaussian5_batch(srcs, dsts, n, 0);
}
int rcv_yuyv_to_bgr_gaussian5_batch_multi(const RcvMat *srcs, RcvMat *dsts, int32_t n, int32_t ngpus) {
  return yuyv_to_bgr_gaussian5_batch(srcs, dsts, n, ngpus > 0 ? ngpus : -1);
}

static int check_yuyv_sobel(const RcvMat *src, const RcvMat *mag) {
  RCV_TRY(check_mat(src, "src"));
  RCV_TRY(check_mat(mag, "mag"));
  if (src->depth != RCV_U8 || src->channels != 2) return fail(RCV_ERR_DEPTH, "YUYV source must be u8 with channels = 2");
  if (mag->depth != RCV_F32 || mag->channels != 1) return fail(RCV_ERR_DEPTH, "magnitude must be f32 with channels = 1");
  if (src->cols & 1) return fail(RCV_ERR_SIZE, "YUYV width %d is odd", src->cols);
  return check_same_size(src, mag, "YUYV->Sobel");
}

int rcv_yuyv_to_sobel_mag(const RcvMat *src, RcvMat *mag) {
  RCV_TRY(check_yuyv_sobel(src, mag));
  return run_unary(src, mag, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_yuyv_sobel(c, s, d, st);
  }, -1, band_window(1));
}

static int yuyv_to_sobel_mag_batch(const RcvMat *srcs, RcvMat *mags, int32_t n, int ngpus) {
  RCV_TRY(check_batch(srcs, mags, n, "YUYV->Sobel"));
  if (n == 0) return RCV_OK;
  RCV_TRY(check_yuyv_sobel(&srcs[0], &mags[0]));
  return run_batch(srcs, mags, n, [](Ctx *c, const DBatch &s, const DBatch &d, cudaStream_t st) {
    return launch_yuyv_sobel(c, s, d, st);
  }, -1, band_window(1), ngpus);
}
int rcv_yuyv_to_sobel_mag_batch(const RcvMat *srcs, RcvMat *mags, int32_t n) { return yuyv_to_sobel_mag_batch(srcs, mags, n, 0); }
int rcv_yuyv_to_sobel_mag_batch_multi(const RcvMat *srcs, RcvMat *mags, int32_t n, int32_t ngpus) {
  return yuyv_to_sobel_mag_batch(srcs, mags, n, ngpus > 0 ? ngpus : -1);
}

// ---- MJPEG branch of read() (videoio/mod.rs:205-232) through nvJPEG ----------------------------------
int rcv_mjpeg_info(const uint8_t *jpeg, size_t len, int32_t *width, int32_t *height) {
  if (!jpeg || !width || !height) return fail(RCV_ERR_ARG, "NULL argument");
  if (len < 4) return fail(RCV_ERR_SIZE, "JPEG buffer of %zu bytes", len);
  Ctx *c = ctx_default();
  if (!c) return RCV_ERR_NOT_INIT;
  std::lock_guard<std::mutex> lk(c->mu);
  RCV_CUDA(cudaSetDevice(c->device));
  int w = 0, h = 0;
  RCV_TRY(mjpeg_info(c, jpeg, len, &w, &h));
  *width = w;
  *height = h;
  return RCV_OK;
}

int rcv_mjpeg_to_bgr(const uint8_t *jpeg, size_t len, RcvMat *dst) {
  if (!jpeg) return fail(RCV_ERR_ARG, "jpeg is NULL");
  if (len < 4) return fail(RCV_ERR_SIZE, "JPEG buffer of %zu bytes", len);
  RCV_TRY(check_mat(dst, "dst"));
  if (dst->depth != RCV_U8 || dst->channels != 3) return fail(RCV_ERR_DEPTH, "MJPEG decodes to u8 with channels = 3");
  const RcvMat *mats[1] = {dst};
  Ctx *c = pick_ctx(mats, 1);
  if (!c) return RCV_ERR_NOT_INIT;
  std::lock_guard<std::mutex> lk(c->mu);
  RCV_CUDA(cudaSetDevice(c->device));
  int w = 0, h = 0;
  RCV_TRY(mjpeg_info(c, jpeg, len, &w, &h));
  if (w != dst->cols || h != dst->rows)
    return fail(RCV_ERR_SIZE, "MJPEG frame is %dx%d, dst is %dx%d (the caller sizes dst: rcv_mjpeg_info)", w, h, dst->cols,
                dst->rows);
  if (w == 0 || h == 0) return RCV_OK;
  Staged so;
  RCV_TRY(stage_alloc(c, dst, SCR_STAGE_OUT0, &so));
  RCV_TRY(launch_mjpeg(c, jpeg, len, so.v, c->stream));
  RCV_TRY(copy_out(dst, so, c->stream));
  if (ctx_blocking() || so.staged) RCV_CUDA(cudaStreamSynchronize(c->stream));
  return RCV_OK;
}

}  // extern "C"
