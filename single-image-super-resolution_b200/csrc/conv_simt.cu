// CUDA-core convolution kernels (NHWC bf16, fp32 accumulate) for the layer shapes the tcgen05
// engine does not take (Cin = 3 or Cout = 3 edge layers, 9x9 first conv) and as the on-device
// cross-check of the tensor-core path in the tests.  Any kernel size / stride / padding.
#include "conv_simt.h"

#include <stdio.h>

namespace sisr {

namespace {

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  if (act == ACT_NONE) return x;
  if (act == ACT_TANH) return tanhf(x);
  return x > 0.f ? x : x * slope;
}

// one thread per (output pixel, output channel); channel fastest.
__global__ void conv_fprop_simt_kernel(SimtConv c, const __nv_bfloat16* __restrict__ x,
                                       const __nv_bfloat16* __restrict__ w,
                                       const float* __restrict__ bias, int act, float slope,
                                       const float* __restrict__ slope_ptr,
                                       __nv_bfloat16* __restrict__ y_bf16,
                                       float* __restrict__ y_nchw_f32) {
  const long long total = static_cast<long long>(c.N) * c.OH * c.OW * c.Cout;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % c.Cout);
  long long pix = idx / c.Cout;
  const int ow = static_cast<int>(pix % c.OW);
  pix /= c.OW;
  const int oh = static_cast<int>(pix % c.OH);
  const int n = static_cast<int>(pix / c.OH);
  float acc = bias ? bias[co] : 0.f;
  const __nv_bfloat16* wr = w + static_cast<size_t>(co) * c.KH * c.KW * c.Cin;
  for (int kh = 0; kh < c.KH; ++kh) {
    const int ih = oh * c.stride - c.pad + kh;
    if (ih < 0 || ih >= c.H) continue;
    for (int kw = 0; kw < c.KW; ++kw) {
      const int iw = ow * c.stride - c.pad + kw;
      if (iw < 0 || iw >= c.W) continue;
      const __nv_bfloat16* xp = x + (static_cast<size_t>(n * c.H + ih) * c.W + iw) * c.Cin;
      const __nv_bfloat16* wp = wr + (kh * c.KW + kw) * c.Cin;
      for (int ci = 0; ci < c.Cin; ++ci)
        acc = fmaf(__bfloat162float(xp[ci]), __bfloat162float(wp[ci]), acc);
    }
  }
  if (act == ACT_PRELU) slope = *slope_ptr;
  if (act == ACT_RELU) slope = 0.f;
  acc = apply_act(acc, act, slope);
  if (y_bf16)
    y_bf16[(static_cast<size_t>(n * c.OH + oh) * c.OW + ow) * c.Cout + co] = __float2bfloat16_rn(acc);
  if (y_nchw_f32)
    y_nchw_f32[(static_cast<size_t>(n * c.Cout + co) * c.OH + oh) * c.OW + ow] = acc;
}

// dx[n,ih,iw,ci] = sum_{kh,kw,co} dy[n,oh,ow,co] * w[co,kh,kw,ci],  oh*stride - pad + kh = ih
__global__ void conv_dgrad_simt_kernel(SimtConv c, const __nv_bfloat16* __restrict__ dy,
                                       const __nv_bfloat16* __restrict__ w,
                                       __nv_bfloat16* __restrict__ dx) {
  const long long total = static_cast<long long>(c.N) * c.H * c.W * c.Cin;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int ci = static_cast<int>(idx % c.Cin);
  long long pix = idx / c.Cin;
  const int iw = static_cast<int>(pix % c.W);
  pix /= c.W;
  const int ih = static_cast<int>(pix % c.H);
  const int n = static_cast<int>(pix / c.H);
  float acc = 0.f;
  for (int kh = 0; kh < c.KH; ++kh) {
    const int th = ih + c.pad - kh;
    if (th < 0 || th % c.stride) continue;
    const int oh = th / c.stride;
    if (oh >= c.OH) continue;
    for (int kw = 0; kw < c.KW; ++kw) {
      const int tw = iw + c.pad - kw;
      if (tw < 0 || tw % c.stride) continue;
      const int ow = tw / c.stride;
      if (ow >= c.OW) continue;
      const __nv_bfloat16* dyp = dy + (static_cast<size_t>(n * c.OH + oh) * c.OW + ow) * c.Cout;
      const __nv_bfloat16* wp = w + static_cast<size_t>(kh * c.KW + kw) * c.Cin + ci;
      const size_t wstride = static_cast<size_t>(c.KH) * c.KW * c.Cin;
      for (int co = 0; co < c.Cout; ++co)
        acc = fmaf(__bfloat162float(dyp[co]), __bfloat162float(wp[co * wstride]), acc);
    }
  }
  dx[idx] = __float2bfloat16_rn(acc);
}

// dw[co,kh,kw,ci] = sum_pixels dy[p,co] * x[p shifted by tap, ci].  One block per (co, tap);
// threads = groups x channel-chunk, block reduction over the pixel groups.
constexpr int kWgradThreads = 256;
__device__ __forceinline__ size_t dy_index(const SimtConv& c, int n, int oh, int ow, int co, int ps_c) {
  if (ps_c == 0) return (static_cast<size_t>(n * c.OH + oh) * c.OW + ow) * c.Cout + co;
  const int sub = co / ps_c, ch = co - sub * ps_c;
  return (static_cast<size_t>(n * 2 * c.OH + 2 * oh + (sub >> 1)) * (2 * c.OW) + 2 * ow + (sub & 1)) *
             ps_c + ch;
}
__global__ void conv_wgrad_simt_kernel(SimtConv c, const __nv_bfloat16* __restrict__ x,
                                       const __nv_bfloat16* __restrict__ dy,
                                       float* __restrict__ dw, float* __restrict__ dbias, int ps_c,
                                       int accumulate) {
  __shared__ float red[kWgradThreads];
  const int co = blockIdx.x;
  const int tap = blockIdx.y;
  const int kh = tap / c.KW, kw = tap % c.KW;
  const int cchunk = c.Cin < kWgradThreads ? c.Cin : kWgradThreads;
  const int groups = kWgradThreads / cchunk;
  const int g = threadIdx.x / cchunk;
  const int cl = threadIdx.x % cchunk;
  const long long npix = static_cast<long long>(c.N) * c.OH * c.OW;
  for (int c0 = 0; c0 < c.Cin; c0 += cchunk) {
    const int ci = c0 + cl;
    float acc = 0.f;
    if (g < groups && ci < c.Cin) {
      for (long long p = g; p < npix; p += groups) {
        const int ow = static_cast<int>(p % c.OW);
        const long long t = p / c.OW;
        const int oh = static_cast<int>(t % c.OH);
        const int n = static_cast<int>(t / c.OH);
        const int ih = oh * c.stride - c.pad + kh;
        const int iw = ow * c.stride - c.pad + kw;
        if (ih < 0 || ih >= c.H || iw < 0 || iw >= c.W) continue;
        const float d = __bfloat162float(dy[dy_index(c, n, oh, ow, co, ps_c)]);
        const float v =
            __bfloat162float(x[(static_cast<size_t>(n * c.H + ih) * c.W + iw) * c.Cin + ci]);
        acc = fmaf(d, v, acc);
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0 && ci < c.Cin) {
      float s = 0.f;
      for (int j = 0; j < groups; ++j) s += red[j * cchunk + cl];
      float* dst = dw + (static_cast<size_t>(co) * c.KH * c.KW + tap) * c.Cin + ci;
      *dst = accumulate ? *dst + s : s;
    }
    __syncthreads();
  }
  if (dbias && tap == 0) {
    float acc = 0.f;
    for (long long p = threadIdx.x; p < npix; p += blockDim.x) {
      const int ow = static_cast<int>(p % c.OW);
      const long long t = p / c.OW;
      acc += __bfloat162float(dy[dy_index(c, static_cast<int>(t / c.OH), static_cast<int>(t % c.OH),
                                          ow, co, ps_c)]);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int j = 0; j < kWgradThreads; ++j) s += red[j];
      dbias[co] = accumulate ? dbias[co] + s : s;
    }
  }
}

inline int blocks_for(long long total, int threads) {
  return static_cast<int>((total + threads - 1) / threads);
}

}  // namespace

int conv_fprop_simt(const SimtConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w,
                    const float* bias, int act, float slope, const float* slope_ptr,
                    __nv_bfloat16* y_bf16, float* y_nchw_f32, cudaStream_t stream) {
  const long long total = static_cast<long long>(c.N) * c.OH * c.OW * c.Cout;
  if (total == 0) return 0;
  conv_fprop_simt_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(c, x, w, bias, act, slope,
                                                                      slope_ptr, y_bf16, y_nchw_f32);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

int conv_dgrad_simt(const SimtConv& c, const __nv_bfloat16* dy, const __nv_bfloat16* w,
                    __nv_bfloat16* dx, cudaStream_t stream) {
  const long long total = static_cast<long long>(c.N) * c.H * c.W * c.Cin;
  if (total == 0) return 0;
  conv_dgrad_simt_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(c, dy, w, dx);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

int conv_wgrad_simt(const SimtConv& c, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
                    float* dbias, int ps_c, int accumulate, cudaStream_t stream) {
  dim3 grid(c.Cout, c.KH * c.KW);
  conv_wgrad_simt_kernel<<<grid, kWgradThreads, 0, stream>>>(c, x, dy, dw, dbias, ps_c, accumulate);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

}  // namespace sisr
