// Internal C++ interface of the discriminator head (see linear.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace sisr {

// x_flat: [B, fc_in] bf16 in the reference's (c,h,w) flatten order; w0: [fc_mid, fc_in] fp32.
// h: [B, fc_mid] fp32 (post-LeakyReLU, kept for backward); p: [B] sigmoid output.
int dhead_forward(const __nv_bfloat16* x_flat, const float* w0, const float* b0, const float* w2,
                  const float* b2, float slope, float* h, float* p, int B, int fc_in, int fc_mid,
                  cudaStream_t s);
// dp: [B] gradient w.r.t. p.  dh: [B, fc_mid] scratch.  dx_flat: [B, fc_in] fp32 or null.
int dhead_backward(const __nv_bfloat16* x_flat, const float* w0, const float* w2, const float* h,
                   const float* p, const float* dp, float slope, float* dh, float* dw0, float* db0,
                   float* dw2, float* db2, float* dx_flat, int B, int fc_in, int fc_mid,
                   int need_wgrad, cudaStream_t s);

}  // namespace sisr
