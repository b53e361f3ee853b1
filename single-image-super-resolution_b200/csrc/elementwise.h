// Internal C++ interface of the bandwidth-bound kernels (see elementwise.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "igemm.h"
#include "peer.h"

namespace sisr {

int nchw_f32_to_nhwc_bf16(const float* x, __nv_bfloat16* y, int N, int C, int H, int W, cudaStream_t s);
int nhwc_bf16_to_nchw_f32(const __nv_bfloat16* x, float* y, int N, int C, int H, int W, cudaStream_t s);
int tanh_bwd_nchw_to_nhwc(const float* dout, const float* y, __nv_bfloat16* dpre, int N, int C, int H,
                          int W, cudaStream_t s);
// x: [batch][R][Cc] -> y: [batch][Cc][R]
int transpose_bf16(const __nv_bfloat16* x, __nv_bfloat16* y, int batch, int R, int Cc, cudaStream_t s);

int col_stats(const __nv_bfloat16* y, long long M, int C, float* stats, int with_sq, cudaStream_t s);
int bn_finalize(const float* stats, int stats_rows, float count, const float* gamma, const float* beta,
                float* running_mean, float* running_var, long long* num_batches, float momentum,
                float eps, int training, float* scale, float* shift, float* mean, float* invstd, int C,
                cudaStream_t s);
int bn_apply(const __nv_bfloat16* y, const float* scale, const float* shift, int act, float slope,
             const float* slope_ptr, const __nv_bfloat16* residual, __nv_bfloat16* out, long long M,
             int C, cudaStream_t s);
int bn_bwd_reduce(const __nv_bfloat16* dout, const __nv_bfloat16* y, const float* mean,
                  const float* invstd, const float* scale, const float* shift, int act, float slope,
                  const float* slope_ptr, float* sums, long long M, int C, cudaStream_t s,
                  const struct PeerTable* peer = nullptr, int slot = 0, float* sums_global = nullptr,
                  unsigned int* ticket = nullptr);
int bn_bwd_apply(const __nv_bfloat16* dout, const __nv_bfloat16* y, const float* mean,
                 const float* invstd, const float* scale, const float* shift, int act, float slope,
                 const float* slope_ptr, const float* sums, float count, __nv_bfloat16* dy,
                 float* colsum, long long M, int C, cudaStream_t s);
int act_bwd(const __nv_bfloat16* dout, const __nv_bfloat16* out, int act, float slope,
            const float* slope_ptr, __nv_bfloat16* din, float* dslope, float* colsum, long long M, int C,
            cudaStream_t s);
int maxpool2_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, int N, int H, int W, int C, cudaStream_t s);
int maxpool2_bwd(const __nv_bfloat16* x, const __nv_bfloat16* dy, __nv_bfloat16* dx, int N, int H,
                 int W, int C, cudaStream_t s);
// PixelShuffle(2) on NHWC bf16: [N,H,W,4C] -> [N,2H,2W,C] (inverse = 1: the gradient mapping)
int pixel_shuffle2(const __nv_bfloat16* src, __nv_bfloat16* dst, int N, int H, int W, int C, int inverse,
                   cudaStream_t s);
// lr = clamp(bicubic(hr)) (dlr == nullptr) or its backward dhr (dlr != nullptr)
int lr_from_hr(const float* hr, float* lr, const float* dlr, float* dhr, int N, int C, int H, int W, int OH,
               int OW, cudaStream_t s);
int mse_fwd(const float* a, const float* b, long long n, float coef, float* loss, cudaStream_t s);
int mse_bwd(const float* a, const float* b, long long n, float coef, const float* gout, float* ga,
            float* gb, cudaStream_t s);
int bce_fwd(const float* p, int n, float target, float* loss, float* mean_p, cudaStream_t s);
int bce_bwd(const float* p, int n, float target, const float* gout, float* dp, cudaStream_t s);

}  // namespace sisr
