// Internal C++ interface of the thin (3-channel) edge-layer kernels (see conv_thin.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "igemm.h"

namespace sisr {

struct ThinConv {
  int N, H, W;   // spatial size (stride 1, same-size convolution)
  int CS, CW;    // small (3) and wide channel counts
  int k, pad;
};

bool thin_in_supported(const ThinConv& c);
bool thin_out_supported(const ThinConv& c);
bool thin_wgrad_supported(const ThinConv& c);
// y[N,H,W,CW] = act(conv(x[N,H,W,CS], w[CW][k*k][CS]) + bias); flip = transposed conv (dgrad)
int thin_in_conv(const ThinConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias,
                 int act, float slope, const float* slope_ptr, int flip, __nv_bfloat16* y, cudaStream_t s);
// y[N,H,W,CS] (bf16) and/or y_nchw[N,CS,H,W] (fp32) = act(conv(x[N,H,W,CW], w[CS][k*k][CW]) + bias)
int thin_out_conv(const ThinConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias,
                  int act, float slope, int flip, __nv_bfloat16* y_bf16, float* y_nchw, cudaStream_t s);
// sgn=+1: out[cw][tap][cs] = sum_q wide[q,cw]*small[q+(tap-pad),cs]   (weight grad of a thin-in conv)
// sgn=-1: out[cs][tap][cw] = sum_q wide[q,cw]*small[q-(tap-pad),cs]   (weight grad of a thin-out conv)
// small_colsum (nullable): [CS] per-channel sum of `small`; wide_colsum (nullable): [CW] per-channel
// sum of `wide` (both overwritten)
int thin_wgrad(const ThinConv& c, const __nv_bfloat16* small, const __nv_bfloat16* wide, int sgn,
               float* out, float* small_colsum, float* wide_colsum, cudaStream_t s);
const char* thin_last_error();

}  // namespace sisr
