// Internal C++ interface of the fused multi-tensor Adam (see optim.cu).
#pragma once
#include <cuda_runtime.h>

namespace sisr {

// Advances the device-side step counter and writes hyper = {lr_t, 1-b1^t, 1-b2^t, t} where
// lr_t = lr0 * decay^(t-1) (LambdaLR stepped once per iteration, config.py:170-180).
int adam_tick(int* step, float lr0, float decay, float b1, float b2, float* hyper, cudaStream_t s);
// torch.optim.Adam update (no weight decay, no amsgrad) over n tensors; grad_scale multiplies g.
int adam_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
               const long long* numel, const float* hyper, float b1, float b2, float eps,
               float grad_scale, cudaStream_t s);

}  // namespace sisr
