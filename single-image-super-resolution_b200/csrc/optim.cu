// Fused multi-tensor Adam: one launch updates up to 24 parameter tensors (pointer table in the
// kernel parameters); hyper-parameters come from device memory so that a captured CUDA graph
// replays with the current step's learning rate and bias corrections.
#include "optim.h"

#include <math.h>
#include <stdint.h>

namespace sisr {

namespace {

constexpr int kMaxTensors = 24;
struct AdamTable {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  long long numel[kMaxTensors];
};

__global__ void adam_tick_kernel(int* step, float lr0, float decay, float b1, float b2, float* hyper) {
  const int t = *step + 1;
  *step = t;
  hyper[0] = lr0 * powf(decay, static_cast<float>(t - 1));
  hyper[1] = 1.f - powf(b1, static_cast<float>(t));
  hyper[2] = 1.f - powf(b2, static_cast<float>(t));
  hyper[3] = static_cast<float>(t);
}

__global__ void adam_multi_kernel(AdamTable tab, const float* __restrict__ hyper, float b1, float b2,
                                  float eps, float grad_scale) {
  const int ti = blockIdx.y;
  const long long n = tab.numel[ti];
  float* __restrict__ p = tab.p[ti];
  const float* __restrict__ g = tab.g[ti];
  float* __restrict__ m = tab.m[ti];
  float* __restrict__ v = tab.v[ti];
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= grad_scale;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  };
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = tid; i < n4; i += nthreads) {       // 16-byte accesses on the bulk of the tensor
    float4 p4 = reinterpret_cast<float4*>(p)[i];
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    upd(p4.x, g4.x, m4.x, v4.x);
    upd(p4.y, g4.y, m4.y, v4.y);
    upd(p4.z, g4.z, m4.z, v4.z);
    upd(p4.w, g4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (long long i = n4 * 4 + tid; i < n; i += nthreads) upd(p[i], g[i], m[i], v[i]);
}

}  // namespace

int adam_tick(int* step, float lr0, float decay, float b1, float b2, float* hyper, cudaStream_t s) {
  adam_tick_kernel<<<1, 1, 0, s>>>(step, lr0, decay, b1, b2, hyper);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

int adam_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
               const long long* numel, const float* hyper, float b1, float b2, float eps,
               float grad_scale, cudaStream_t s) {
  for (int base = 0; base < n; base += kMaxTensors) {
    AdamTable tab;
    const int cnt = n - base < kMaxTensors ? n - base : kMaxTensors;
    long long biggest = 1;
    for (int i = 0; i < cnt; ++i) {
      tab.p[i] = p[base + i];
      tab.g[i] = g[base + i];
      tab.m[i] = m[base + i];
      tab.v[i] = v[base + i];
      tab.numel[i] = numel[base + i];
      if (numel[base + i] > biggest) biggest = numel[base + i];
    }
    long long bx = (biggest + 256 * 8 - 1) / (256 * 8);
    if (bx > 148 * 8) bx = 148 * 8;
    dim3 grid(static_cast<unsigned>(bx), cnt);
    adam_multi_kernel<<<grid, 256, 0, s>>>(tab, hyper, b1, b2, eps, grad_scale);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

}  // namespace sisr
