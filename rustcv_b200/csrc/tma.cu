// tma.cu -- host-side tensor-map encoding for the strip-pipeline kernels.
//
// cuTensorMapEncodeTiled is a driver-API function; it is fetched through
// cudaGetDriverEntryPoint so the library links against cudart only.
#include "rcv_internal.cuh"

#include <mutex>

namespace rcv {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;

static void load_encode() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_encode = (EncodeTiledFn)fn;
}

int make_tmap_rows_u32(CUtensorMap *out, const void *base, size_t row_bytes, int rows, size_t step, int n,
                       size_t frame_stride, int box_w_words, int box_h) {
  std::call_once(g_once, load_encode);
  if (!g_encode) return fail(RCV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if (((uintptr_t)base & 15) || (step & 15) || (frame_stride & 15))
    return fail(RCV_ERR_ARG, "TMA needs 16-byte aligned base/step/frame stride");
  cuuint64_t dims[3] = {(cuuint64_t)((row_bytes + 3) / 4), (cuuint64_t)rows, (cuuint64_t)(n < 1 ? 1 : n)};
  // a 1-frame batch still needs a legal (multiple of 16, non-zero) stride for dim 2
  cuuint64_t fs = (n > 1 && frame_stride) ? frame_stride : (cuuint64_t)step * rows;
  cuuint64_t strides[2] = {(cuuint64_t)step, fs};
  cuuint32_t box[3] = {(cuuint32_t)box_w_words, (cuuint32_t)box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RCV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return RCV_OK;
}

}  // namespace rcv
