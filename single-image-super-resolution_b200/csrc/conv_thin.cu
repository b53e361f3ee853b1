// CUDA-core kernels for the "thin" edge layers of the SRGAN step, where one side of the
// convolution has 3 channels (RGB) and tensor-core tiles would be > 90 % padding:
//   thin-in  : C_small -> C_wide  (G first conv 9x9, D / VGG first conv 3x3, dgrad of the G output conv)
//   thin-out : C_wide  -> C_small (G output conv 3x3 + Tanh, dgrad of the D / VGG first conv)
//   thin-wgrad: correlation of a 3-channel tensor with a wide tensor over the k x k taps
// Stride 1 only.  NHWC bf16 activations, fp32 accumulate, weights staged in shared memory as fp32.
// These layers are < 2 % of the step's FLOPs; the goal is to keep them near their memory floor.
#include "conv_thin.h"

#include <stdio.h>

#include "ptx.cuh"

namespace sisr {

namespace {

__device__ __forceinline__ float act_apply(float x, int act, float slope) {
  if (act == ACT_NONE) return x;
  if (act == ACT_TANH) return tanhf(x);
  return x > 0.f ? x : x * slope;
}

// ------------------------------------------------------------------ thin-in: CS -> CW, k x k
// w: [CW][k*k][CS] bf16 (prepared layout); flip: use tap (k*k-1-t) (transposed conv)
template <int CS>
__global__ void __launch_bounds__(128)
thin_in_kernel(ThinConv c, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
               const float* __restrict__ bias, int act, float slope, const float* __restrict__ slope_ptr,
               int flip, __nv_bfloat16* __restrict__ y) {
  extern __shared__ float s_w[];  // [(tap*CS + cs)][CW]
  const int T = c.k * c.k;
  for (int i = threadIdx.x; i < c.CW * T * CS; i += blockDim.x) {
    const int cs = i % CS, tap = (i / CS) % T, cw = i / (CS * T);
    const int tt = flip ? T - 1 - tap : tap;
    s_w[(tt * CS + cs) * c.CW + cw] = __bfloat162float(w[i]);
  }
  __syncthreads();
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= npix) return;
  const int ow = static_cast<int>(p % c.W);
  const int oh = static_cast<int>((p / c.W) % c.H);
  const int n = static_cast<int>(p / (static_cast<long long>(c.W) * c.H));
  if (act == ACT_PRELU) slope = *slope_ptr;
  if (act == ACT_RELU) slope = 0.f;
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * c.H * c.W * CS;
  __nv_bfloat16* yp = y + p * c.CW;
  for (int c0 = 0; c0 < c.CW; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = bias ? bias[c0 + j] : 0.f;
    for (int kh = 0; kh < c.k; ++kh) {
      const int ih = oh - c.pad + kh;
      if (ih < 0 || ih >= c.H) continue;
      for (int kw = 0; kw < c.k; ++kw) {
        const int iw = ow - c.pad + kw;
        if (iw < 0 || iw >= c.W) continue;
        const __nv_bfloat16* xp = xn + (static_cast<size_t>(ih) * c.W + iw) * CS;
        const float* wr = s_w + ((kh * c.k + kw) * CS) * c.CW + c0;
#pragma unroll
        for (int cs = 0; cs < CS; ++cs) {
          const float xv = __bfloat162float(xp[cs]);
          const float4* w4 = reinterpret_cast<const float4*>(wr + cs * c.CW);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = w4[q];
            acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
          }
        }
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      pk[j] = pack_bf16x2(act_apply(acc[2 * j], act, slope), act_apply(acc[2 * j + 1], act, slope));
    uint4* d4 = reinterpret_cast<uint4*>(yp + c0);
    d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// ------------------------------------------------------------------ thin-out: CW -> CS, k x k
// w: [CS][k*k][CW] bf16; y_bf16: [N,H,W,CS] and/or y_nchw: [N,CS,H,W] fp32
template <int CS>
__global__ void __launch_bounds__(128)
thin_out_kernel(ThinConv c, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                const float* __restrict__ bias, int act, float slope, int flip,
                __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_nchw) {
  extern __shared__ float s_w[];  // [tap][cw][4]
  const int T = c.k * c.k;
  for (int i = threadIdx.x; i < T * c.CW * 4; i += blockDim.x) {
    const int cs = i & 3, cw = (i >> 2) % c.CW, tap = i / (4 * c.CW);
    const int tt = flip ? T - 1 - tap : tap;
    s_w[i] = cs < CS ? __bfloat162float(w[(static_cast<size_t>(cs) * T + tt) * c.CW + cw]) : 0.f;
  }
  __syncthreads();
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= npix) return;
  const int ow = static_cast<int>(p % c.W);
  const int oh = static_cast<int>((p / c.W) % c.H);
  const int n = static_cast<int>(p / (static_cast<long long>(c.W) * c.H));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int cs = 0; cs < CS; ++cs) acc[cs] = bias[cs];
  }
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * c.H * c.W * c.CW;
  for (int kh = 0; kh < c.k; ++kh) {
    const int ih = oh - c.pad + kh;
    if (ih < 0 || ih >= c.H) continue;
    for (int kw = 0; kw < c.k; ++kw) {
      const int iw = ow - c.pad + kw;
      if (iw < 0 || iw >= c.W) continue;
      const uint4* xp = reinterpret_cast<const uint4*>(xn + (static_cast<size_t>(ih) * c.W + iw) * c.CW);
      const float4* wt = reinterpret_cast<const float4*>(s_w) + (kh * c.k + kw) * c.CW;
      for (int v = 0; v < c.CW / 8; ++v) {
        const uint4 u = xp[v];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          const float4 w0 = wt[v * 8 + 2 * j], w1 = wt[v * 8 + 2 * j + 1];
          acc[0] = fmaf(f.x, w0.x, acc[0]); acc[1] = fmaf(f.x, w0.y, acc[1]);
          acc[2] = fmaf(f.x, w0.z, acc[2]);
          acc[0] = fmaf(f.y, w1.x, acc[0]); acc[1] = fmaf(f.y, w1.y, acc[1]);
          acc[2] = fmaf(f.y, w1.z, acc[2]);
          if (CS == 4) { acc[3] = fmaf(f.x, w0.w, acc[3]); acc[3] = fmaf(f.y, w1.w, acc[3]); }
        }
      }
    }
  }
#pragma unroll
  for (int cs = 0; cs < CS; ++cs) {
    const float v = act_apply(acc[cs], act, slope);
    if (y_bf16) y_bf16[p * CS + cs] = __float2bfloat16_rn(v);
    if (y_nchw) y_nchw[((static_cast<size_t>(n) * CS + cs) * c.H + oh) * c.W + ow] = v;
  }
}

// ------------------------------------------------------------------ thin wgrad
// G[cs, tap, cw] = sum_q Wt[q, cw] * S[q + sgn*(tap - pad), cs]
//   type A (sgn=+1): S = input x (CS ch), Wt = dy (CW ch): out[(cw*T + tap)*CS + cs]
//   type B (sgn=-1): S = dy (CS ch), Wt = input x (CW ch): out[(cs*T + tap)*CW + cw]
// Block = 16 channel-quads x G combo groups x PZ pixels in flight; fp32 atomics into `out`.
template <int CS, int KK, int G>
__global__ void __launch_bounds__(256)
thin_wgrad_kernel(ThinConv c, const __nv_bfloat16* __restrict__ s, const __nv_bfloat16* __restrict__ wt,
                  int sgn, float* __restrict__ out) {
  constexpr int PZ = 256 / (16 * G);
  const int T = c.k * c.k;
  const int combos = T * CS;
  const int cq = threadIdx.x % 16;
  const int g = (threadIdx.x / 16) % G;
  const int pz = threadIdx.x / (16 * G);
  const int cw0 = blockIdx.y * 64 + cq * 4;
  int dh[KK], dw[KK], cs_[KK];
  bool on[KK];
#pragma unroll
  for (int i = 0; i < KK; ++i) {
    const int j = g + G * i;
    on[i] = j < combos;
    const int tap = on[i] ? j / CS : 0;
    cs_[i] = on[i] ? j % CS : 0;
    dh[i] = sgn * (tap / c.k - c.pad);
    dw[i] = sgn * (tap % c.k - c.pad);
  }
  float acc[KK][4];
#pragma unroll
  for (int i = 0; i < KK; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long q0 = blockIdx.x * per;
  const long long q1 = q0 + per < npix ? q0 + per : npix;
  for (long long q = q0 + pz; q < q1; q += PZ) {
    const int qw = static_cast<int>(q % c.W);
    const int qh = static_cast<int>((q / c.W) % c.H);
    const long long nbase = q - (static_cast<long long>(qh) * c.W + qw);   // first pixel of the image
    const uint2 u = *reinterpret_cast<const uint2*>(wt + q * c.CW + cw0);
    const float2 f01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 f23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
#pragma unroll
    for (int i = 0; i < KK; ++i) {
      const int ih = qh + dh[i], iw = qw + dw[i];
      if (on[i] && ih >= 0 && ih < c.H && iw >= 0 && iw < c.W) {
        const float sv = __bfloat162float(s[(nbase + static_cast<long long>(ih) * c.W + iw) * CS + cs_[i]]);
        acc[i][0] = fmaf(sv, f01.x, acc[i][0]);
        acc[i][1] = fmaf(sv, f01.y, acc[i][1]);
        acc[i][2] = fmaf(sv, f23.x, acc[i][2]);
        acc[i][3] = fmaf(sv, f23.y, acc[i][3]);
      }
    }
  }
  // combine the PZ pixel groups in shared memory, then one global atomic per (combo, channel)
  constexpr int kRed = (PZ > 1) ? KK * G * 64 : 1;
  __shared__ float s_red[kRed];
  if (PZ > 1) {
    for (int i = threadIdx.x; i < kRed; i += blockDim.x) s_red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KK; ++i) {
      if (!on[i]) continue;
      const int j = g + G * i;
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&s_red[j * 64 + cq * 4 + e], acc[i][e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < combos * 64; i += blockDim.x) {
      const int j = i / 64, cl = i % 64;
      const int tap = j / CS, cs = j % CS, cw = blockIdx.y * 64 + cl;
      float* dst = (sgn > 0) ? out + (static_cast<size_t>(cw) * T + tap) * CS + cs
                             : out + (static_cast<size_t>(cs) * T + tap) * c.CW + cw;
      atomicAdd(dst, s_red[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < KK; ++i) {
      if (!on[i]) continue;
      const int j = g + G * i;
      const int tap = j / CS;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cw = cw0 + e;
        float* dst = (sgn > 0) ? out + (static_cast<size_t>(cw) * T + tap) * CS + cs_[i]
                               : out + (static_cast<size_t>(cs_[i]) * T + tap) * c.CW + cw;
        atomicAdd(dst, acc[i][e]);
      }
    }
  }
}

// per-channel sum of a small-channel tensor: out[cs] += sum_q s[q, cs]
template <int CS>
__global__ void thin_colsum_kernel(const __nv_bfloat16* __restrict__ s, long long npix,
                                   float* __restrict__ out) {
  float acc[CS];
#pragma unroll
  for (int i = 0; i < CS; ++i) acc[i] = 0.f;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < npix;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int i = 0; i < CS; ++i) acc[i] += __bfloat162float(s[q * CS + i]);
  }
#pragma unroll
  for (int i = 0; i < CS; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[i], v);
  }
}

thread_local char g_err[256] = "";
int check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
  return 4;
}

}  // namespace

const char* thin_last_error() { return g_err; }

bool thin_in_supported(const ThinConv& c) { return c.CS == 3 && c.CW % 16 == 0 && c.CW * c.k * c.k * c.CS * 4 <= 200 * 1024; }
bool thin_out_supported(const ThinConv& c) { return c.CS == 3 && c.CW % 8 == 0 && c.k * c.k * c.CW * 16 <= 200 * 1024; }
bool thin_wgrad_supported(const ThinConv& c) { return c.CS == 3 && c.CW % 64 == 0 && (c.k == 3 || c.k == 9); }

int thin_in_conv(const ThinConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias,
                 int act, float slope, const float* slope_ptr, int flip, __nv_bfloat16* y,
                 cudaStream_t s) {
  const int smem = c.CW * c.k * c.k * c.CS * sizeof(float);
  static int configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(thin_in_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = smem;
  }
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  thin_in_kernel<3><<<static_cast<int>((npix + 127) / 128), 128, smem, s>>>(c, x, w, bias, act, slope,
                                                                           slope_ptr, flip, y);
  return check("thin_in_conv");
}

int thin_out_conv(const ThinConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias,
                  int act, float slope, int flip, __nv_bfloat16* y_bf16, float* y_nchw, cudaStream_t s) {
  const int smem = c.k * c.k * c.CW * 4 * sizeof(float);
  static int configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(thin_out_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = smem;
  }
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  thin_out_kernel<3><<<static_cast<int>((npix + 127) / 128), 128, smem, s>>>(c, x, w, bias, act, slope,
                                                                            flip, y_bf16, y_nchw);
  return check("thin_out_conv");
}

int thin_wgrad(const ThinConv& c, const __nv_bfloat16* small, const __nv_bfloat16* wide, int sgn,
               float* out, float* small_colsum, cudaStream_t s) {
  const int T = c.k * c.k;
  cudaMemsetAsync(out, 0, sizeof(float) * T * c.CS * c.CW, s);
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  long long gx = (npix + 511) / 512;
  if (gx > 148 * 2) gx = 148 * 2;
  if (gx < 1) gx = 1;
  dim3 grid(static_cast<unsigned>(gx), c.CW / 64);
  if (c.k == 3)
    thin_wgrad_kernel<3, 7, 4><<<grid, 256, 0, s>>>(c, small, wide, sgn, out);
  else
    thin_wgrad_kernel<3, 16, 16><<<grid, 256, 0, s>>>(c, small, wide, sgn, out);
  if (small_colsum) {
    cudaMemsetAsync(small_colsum, 0, sizeof(float) * c.CS, s);
    thin_colsum_kernel<3><<<148, 256, 0, s>>>(small, npix, small_colsum);
  }
  return check("thin_wgrad");
}

}  // namespace sisr
