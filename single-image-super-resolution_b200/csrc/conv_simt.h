// CUDA-core convolution kernels (internal C++ interface).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "igemm.h"

namespace sisr {

struct SimtConv {
  int N, H, W, Cin;     // input NHWC
  int OH, OW, Cout;     // output NHWC
  int KH, KW, stride, pad;
};

// weights are [Cout, KH, KW, Cin] bf16.  Either output pointer may be null.
int conv_fprop_simt(const SimtConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w,
                    const float* bias, int act, float slope, const float* slope_ptr,
                    __nv_bfloat16* y_bf16, float* y_nchw_f32, cudaStream_t stream);
int conv_dgrad_simt(const SimtConv& c, const __nv_bfloat16* dy, const __nv_bfloat16* w,
                    __nv_bfloat16* dx, cudaStream_t stream);
// dw is fp32 [Cout, KH, KW, Cin]; dbias fp32 [Cout] or null.  ps_c > 0: dy is stored
// pixel-shuffled as [N, 2*OH, 2*OW, ps_c] and GEMM column co = (i*2+j)*ps_c + c.
int conv_wgrad_simt(const SimtConv& c, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
                    float* dbias, int ps_c, int accumulate, cudaStream_t stream);

}  // namespace sisr
