#include "tmap.h"

#include <cudaTypedefs.h>
#include <stdio.h>

#include <mutex>

namespace sisr {

namespace {
PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;
std::once_flag g_once;
thread_local char g_err[256] = "";

void load_entry_points() {
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) ==
          cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
}
// The encode entry points are driver API: they need a context current on the CALLING thread.  The
// runtime binds the primary context lazily on a thread's first runtime call, and a PyTorch autograd
// worker thread may reach a conv backward without having made one (cached allocations, no kernel yet).
void bind_context() {
  thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}
}  // namespace

const char* tmap_last_error() { return g_err; }

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows) {
  std::call_once(g_once, load_entry_points);
  bind_context();
  if (!g_encode_tiled) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled entry point unavailable");
    return 1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err,
             "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu stride=%llu box=%ux%u",
             (int)r, (unsigned long long)rows, (unsigned long long)cols,
             (unsigned long long)row_stride_elems, box_cols, box_rows);
    return 2;
  }
  return 0;
}

int make_tmap_tiled_nhwc_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C,
                              uint32_t channels, uint32_t box_w, uint32_t box_h) {
  std::call_once(g_once, load_entry_points);
  bind_context();
  if (!g_encode_tiled) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled entry point unavailable");
    return 1;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * W * 2, (cuuint64_t)C * W * H * 2};
  cuuint32_t box[4] = {channels, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled(4d) failed (%d): NHWC=%d,%d,%d,%d box=%u,%u,%u",
             (int)r, N, H, W, C, channels, box_w, box_h);
    return 2;
  }
  return 0;
}

int make_tmap_im2col_nhwc_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C,
                               int lower_w, int lower_h, int upper_w, int upper_h,
                               uint32_t channels, uint32_t pixels, uint32_t trav_stride) {
  std::call_once(g_once, load_entry_points);
  bind_context();
  if (!g_encode_im2col) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeIm2col entry point unavailable");
    return 1;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * W * 2, (cuuint64_t)C * W * H * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, trav_stride, trav_stride, 1};
  CUresult r = g_encode_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                               dims, strides, lower, upper, channels, pixels, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err,
             "cuTensorMapEncodeIm2col failed (%d): NHWC=%d,%d,%d,%d lower=%d,%d upper=%d,%d ch=%u "
             "px=%u stride=%u",
             (int)r, N, H, W, C, lower_w, lower_h, upper_w, upper_h, channels, pixels, trav_stride);
    return 2;
  }
  return 0;
}

}  // namespace sisr
