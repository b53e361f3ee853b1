// Discriminator head: Linear(fc_in -> fc_mid) + LeakyReLU + Linear(fc_mid -> 1) + Sigmoid,
// forward and backward.  The large GEMMs run on a strided CUDA-core tile kernel that reads the
// fp32 master weight in its native [fc_mid, C*H*W] layout (the activations are permuted to the
// reference's (c,h,w) flatten order instead of permuting the 19 M-element weight).
#include "linear.h"

#include <stdio.h>

namespace sisr {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// C[m,n] (+)= sum_k A(m,k) * B(n,k);  A(m,k) = a[m*sam + k*sak], B(n,k) = b[n*sbn + k*sbk]
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ a, long long sam, long long sak, const TB* __restrict__ b,
                 long long sbn, long long sbk, float* __restrict__ c, long long ldc, int M, int N,
                 int K, int k_per_split) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int t = threadIdx.x;
  const int tx = t % 16, ty = t / 16;
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m, k;
      if (sak == 1) { k = t % TK; m = t / TK + 16 * j; } else { m = t % TM; k = t / TM + 4 * j; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? to_f(a[gm * sam + gk * sak]) : 0.f;
      int n, kk;
      if (sbk == 1) { kk = t % TK; n = t / TK + 16 * j; } else { n = t % TN; kk = t / TN + 4 * j; }
      const int gn = n0 + n, gk2 = k0 + kk;
      Bs[kk][n] = (gn < N && gk2 < kend) ? to_f(b[gn * sbn + gk2 * sbk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* dst = c + gm * ldc + gn;
      if (gridDim.z > 1) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
    }
  }
}

template <typename TA, typename TB>
int launch_gemm(const TA* a, long long sam, long long sak, const TB* b, long long sbn, long long sbk,
                float* c, long long ldc, int M, int N, int K, cudaStream_t s) {
  const int tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN);
  int splits = 1;
  if (tiles < 148 && K >= 1024) {
    splits = (296 + tiles - 1) / tiles;
    const int max_splits = K / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = (K + splits - 1) / splits;
  kps = (kps + TK - 1) / TK * TK;
  splits = (K + kps - 1) / kps;
  if (splits > 1) cudaMemsetAsync(c, 0, sizeof(float) * static_cast<size_t>(M) * ldc, s);
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, splits);
  gemm_simt_kernel<TA, TB><<<grid, 256, 0, s>>>(a, sam, sak, b, sbn, sbk, c, ldc, M, N, K, kps);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

__device__ __forceinline__ float block_sum256(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < 8; ++i) t += scratch[i];
  return t;
}

// h = leaky(h_pre + b0) (in place), p = sigmoid(h . w2 + b2); one block per sample
__global__ void dhead_tail_fwd_kernel(float* __restrict__ h, const float* __restrict__ b0,
                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                      float slope, float* __restrict__ p, int mid) {
  __shared__ float scratch[8];
  float* hb = h + static_cast<size_t>(blockIdx.x) * mid;
  float acc = 0.f;
  for (int j = threadIdx.x; j < mid; j += blockDim.x) {
    float v = hb[j] + b0[j];
    v = v > 0.f ? v : v * slope;
    hb[j] = v;
    acc = fmaf(v, w2[j], acc);
  }
  acc = block_sum256(acc, scratch);
  if (threadIdx.x == 0) p[blockIdx.x] = 1.f / (1.f + expf(-(acc + b2[0])));
}
// dz = dp*p*(1-p); dh_pre = dz*w2*leaky'(h); dw2 += dz*h; db2 += dz; db0 += dh_pre
__global__ void dhead_tail_bwd_kernel(const float* __restrict__ h, const float* __restrict__ w2,
                                      const float* __restrict__ p, const float* __restrict__ dp,
                                      float slope, float* __restrict__ dh, float* __restrict__ dw2,
                                      float* __restrict__ db2, float* __restrict__ db0, int mid) {
  const int bidx = blockIdx.x;
  const float pv = p[bidx];
  const float dz = dp[bidx] * pv * (1.f - pv);
  const float* hb = h + static_cast<size_t>(bidx) * mid;
  float* dhb = dh + static_cast<size_t>(bidx) * mid;
  for (int j = threadIdx.x; j < mid; j += blockDim.x) {
    const float hv = hb[j];
    const float g = dz * w2[j] * (hv > 0.f ? 1.f : slope);
    dhb[j] = g;
    atomicAdd(&db0[j], g);
    atomicAdd(&dw2[j], dz * hv);
  }
  if (threadIdx.x == 0) atomicAdd(db2, dz);
}

}  // namespace

int dhead_forward(const __nv_bfloat16* x_flat, const float* w0, const float* b0, const float* w2,
                  const float* b2, float slope, float* h, float* p, int B, int fc_in, int fc_mid,
                  cudaStream_t s) {
  if (int rc = launch_gemm<__nv_bfloat16, float>(x_flat, fc_in, 1, w0, fc_in, 1, h, fc_mid, B, fc_mid,
                                                 fc_in, s))
    return rc;
  dhead_tail_fwd_kernel<<<B, 256, 0, s>>>(h, b0, w2, b2, slope, p, fc_mid);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

int dhead_backward(const __nv_bfloat16* x_flat, const float* w0, const float* w2, const float* h,
                   const float* p, const float* dp, float slope, float* dh, float* dw0, float* db0,
                   float* dw2, float* db2, float* dx_flat, int B, int fc_in, int fc_mid,
                   int need_wgrad, cudaStream_t s) {
  cudaMemsetAsync(db0, 0, sizeof(float) * fc_mid, s);
  cudaMemsetAsync(dw2, 0, sizeof(float) * fc_mid, s);
  cudaMemsetAsync(db2, 0, sizeof(float), s);
  dhead_tail_bwd_kernel<<<B, 256, 0, s>>>(h, w2, p, dp, slope, dh, dw2, db2, db0, fc_mid);
  if (need_wgrad) {
    // dW0[m=fc_mid, n=fc_in] = sum_b dh[b,m] * x[b,n]
    if (int rc = launch_gemm<float, __nv_bfloat16>(dh, 1, fc_mid, x_flat, 1, fc_in, dw0, fc_in, fc_mid,
                                                   fc_in, B, s))
      return rc;
  }
  if (dx_flat) {
    // dx[b, n=fc_in] = sum_k dh[b,k] * W0[k,n]
    if (int rc = launch_gemm<float, float>(dh, fc_mid, 1, w0, 1, fc_in, dx_flat, fc_in, B, fc_in,
                                           fc_mid, s))
      return rc;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

}  // namespace sisr
