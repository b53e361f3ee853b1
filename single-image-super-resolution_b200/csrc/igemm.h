// Internal C++ interface of the tcgen05 implicit-GEMM engine (not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sisr {

constexpr int kMaxTaps = 36;  // 9 filter taps x 4 PixelShuffle sub-pixels (dgrad through the shuffle)

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_PRELU = 3, ACT_TANH = 4 };

// One "tap" = one (dy, dx) filter position: the im2col offset added to the traversal position
// and the first weight-matrix column holding that tap's Cin channels.
struct IgemmTaps {
  uint16_t off_w[kMaxTaps];
  uint16_t off_h[kMaxTaps];
  int32_t k_off[kMaxTaps];
};

// D[M, Cout] = sum_taps A_tap[M, Cin] * Wt[Cout, k_off(tap) .. +Cin]^T
//   rows of A are the NB*GH*GW positions of a traversal grid over an NHWC bf16 tensor:
//   position (n, gh, gw) reads input pixel (gh*trav_stride + lower_h + off_h,
//                                           gw*trav_stride + lower_w + off_w), zero outside.
struct IgemmProblem {
  // input activation tensor
  const __nv_bfloat16* x;
  int NB, H, W, Cin;
  // traversal grid / bounding box
  int GH, GW;
  int trav_stride;
  int lower_w, lower_h, upper_w, upper_h;
  // weights [Cout, Ktot] row-major bf16 (K contiguous)
  const __nv_bfloat16* w;
  int Cout, Ktot;
  int num_taps;
  IgemmTaps taps;
  // output: row (n, gh, gw), column c ->
  //   out[((n*OH + gh*osy + opy)*OW + gw*osx + opx)*ldc + c]          (ps_c == 0)
  //   out[((n*OH + gh*2 + i)*OW + gw*2 + j)*ldc + (c % ps_c)], (i,j) = divmod(c / ps_c, 2)  (ps_c > 0)
  __nv_bfloat16* out;
  int OH, OW, ldc;
  int osy, osx, opy, opx;
  int ps_c;
  // optional tap classes (stride-2 dgrad: the four output parities in ONE launch): class c uses taps
  // [cls_tap_begin[c], +cls_tap_count[c]) and writes to (gh*osy + cls_opy[c], gw*osx + cls_opx[c]);
  // n_classes <= 1: all num_taps taps, (opy, opx)
  int n_classes;
  int cls_tap_begin[4], cls_tap_count[4], cls_opy[4], cls_opx[4];
  // fused epilogue
  const float* bias;       // [Cout] or nullptr (indexed by GEMM column)
  int act;                 // Act
  float slope;             // ACT_LEAKY
  const float* slope_ptr;  // ACT_PRELU (single learnable slope)
  const __nv_bfloat16* mask;  // nullable, indexed like `out`: out = (mask > 0) ? value : value * mask_slope
  float mask_slope;           // (ReLU backward of the tensor this dgrad feeds, fused into the store)
  float* stats;            // [stats_rows][2*Cout]: row r = {sum, sum of squares} of the stored bf16 values
                           // over the tiles of CTA r (rows >= grid are zeroed), or nullptr
  int stats_rows;          // >= igemm_max_ctas()
};

// Returns 0 on success; message via igemm_last_error().
int igemm_launch(const IgemmProblem& p, cudaStream_t stream);
bool igemm_supported(const IgemmProblem& p);
void igemm_set_transposed(int on);     // 0: never put the pixels on the UMMA N side (A-B timing)
// Halo-fed kernel with the pixels on the UMMA M side (igemm_pm.cu): same-size stride-1 3x3, 64 -> 64 channels.
// igemm_launch() routes to it first when igemm_pm_supported(p).
bool igemm_pm_supported(const IgemmProblem& p);
int igemm_pm_launch(const IgemmProblem& p, cudaStream_t stream);
void igemm_set_pm(int on);             // 0: never use it (A-B timing, SISR_PM=0)
void igemm_set_pm_grp(int grp);
// harness only: 8 device counters that CTA 0 adds its phase cycles to ({MMA warp: wait accumulator, wait box, issue,
// tiles}, {epilogue warp 2: wait M tile, work, M tiles}, kernel cycles); nullptr (default) = no instrumentation
void igemm_set_pm_debug(long long* counters);        // M tiles whose instructions are interleaved (0: planner's choice)
const char* igemm_pm_last_error();
int igemm_max_ctas();   // the persistent grid never exceeds this (number of SMs)
const char* igemm_last_error();

}  // namespace sisr
