// Host-side TMA descriptor construction.  The driver entry points are fetched at run time
// through cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda
// (the build box has no driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sisr {

// 2-D row-major bf16 matrix [rows, cols] (cols contiguous); box = {box_cols, box_rows},
// 128-byte swizzle.  box_cols * 2 bytes must be <= 128.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows);

// im2col-mode descriptor over an NHWC bf16 tensor [N, H, W, C].
//   lower/upper   : bounding-box corner offsets (same value for W and H)
//   channels      : channels per pixel fetched per request (<= 64 with 128 B swizzle)
//   pixels        : pixels per request (GEMM-M tile)
//   trav_stride   : traversal stride of the bounding-box position (= conv stride)
int make_tmap_im2col_nhwc_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C,
                               int lower_w, int lower_h, int upper_w, int upper_h,
                               uint32_t channels, uint32_t pixels, uint32_t trav_stride);

// tiled-mode descriptor over an NHWC bf16 tensor [N, H, W, C]: box = {channels, box_w, box_h, 1},
// 128-byte swizzle, zero fill outside the tensor (negative / overflowing coordinates allowed).
int make_tmap_tiled_nhwc_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C,
                              uint32_t channels, uint32_t box_w, uint32_t box_h);

const char* tmap_last_error();

}  // namespace sisr
