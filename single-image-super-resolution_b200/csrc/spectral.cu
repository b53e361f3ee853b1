// Spectral-norm power iteration (legacy torch.nn.utils.spectral_norm semantics), weight
// preparation (1/sigma scaling, [Cout,Cin,kh,kw] fp32 -> [Cout',kh,kw,Cin] bf16 for fprop and
// [Cin,kh,kw,Cout'] bf16 for dgrad, optional PixelShuffle row permutation) and the mapping of the
// weight gradient back to the fp32 master layout including the gradient through sigma.
#include "spectral.h"

#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

namespace sisr {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += scratch[i];
  return t;
}

// t[k] += sum_{co in chunk} W[co,k] * u[co]
__global__ void sn_wtu_kernel(const float* __restrict__ w, const float* __restrict__ u,
                              float* __restrict__ t, int Cout, int K, int co_chunk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int c0 = blockIdx.y * co_chunk;
  const int c1 = min(Cout, c0 + co_chunk);
  float acc = 0.f;
  for (int co = c0; co < c1; ++co) acc = fmaf(w[static_cast<size_t>(co) * K + k], u[co], acc);
  atomicAdd(&t[k], acc);
}
// s[co] = sum_k W[co,k] * t[k]
__global__ void sn_wv_kernel(const float* __restrict__ w, const float* __restrict__ t,
                             float* __restrict__ s, int K) {
  __shared__ float scratch[32];
  const int co = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    acc = fmaf(w[static_cast<size_t>(co) * K + k], t[k], acc);
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) s[co] = acc;
}
// training: v = t/|t|, s = s_raw/|t|, u = s/|s|, sigma = u.s ; eval: sigma = u . s_raw (t was v)
__global__ void sn_finish_kernel(const float* __restrict__ t, const float* __restrict__ s_raw,
                                 float* __restrict__ u, float* __restrict__ v,
                                 float* __restrict__ sigma, int Cout, int K, int training, float eps) {
  __shared__ float scratch[32];
  if (training) {
    float a = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) a += t[k] * t[k];
    const float nt = fmaxf(sqrtf(block_sum(a, scratch)), eps);
    for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] = t[k] / nt;
    float b = 0.f;
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
      const float sv = s_raw[c] / nt;
      b += sv * sv;
    }
    const float ss = block_sum(b, scratch);
    const float ns = fmaxf(sqrtf(ss), eps);
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) u[c] = (s_raw[c] / nt) / ns;
    if (threadIdx.x == 0) *sigma = ss / ns;
  } else {
    float b = 0.f;
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) b += u[c] * s_raw[c];
    b = block_sum(b, scratch);
    if (threadIdx.x == 0) *sigma = b;
  }
}

__device__ __forceinline__ int unpermute_row(int cop, int Cout, int ps_r) {
  if (ps_r <= 1) return cop;
  const int r2 = ps_r * ps_r;
  const int cps = Cout / r2;
  const int sub = cop / cps, c = cop % cps;
  return c * r2 + sub;  // PixelShuffle: co = c*r^2 + i*r + j, sub = i*r + j
}

__global__ void weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ sigma,
                                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd, float* __restrict__ bias_perm,
                                   int Cout, int Cin, int KH, int KW, int ps_r) {
  const float inv = sigma ? 1.f / *sigma : 1.f;
  const int taps = KH * KW;
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    const int tap = static_cast<int>((i / Cin) % taps);
    const int cop = static_cast<int>(i / (static_cast<long long>(Cin) * taps));
    const int co = unpermute_row(cop, Cout, ps_r);
    const float val = w[(static_cast<size_t>(co) * Cin + ci) * taps + tap] * inv;
    const __nv_bfloat16 h = __float2bfloat16_rn(val);
    wf[i] = h;
    if (wd) wd[(static_cast<size_t>(ci) * taps + tap) * Cout + cop] = h;
    if (bias_perm && tap == 0 && ci == 0) bias_perm[cop] = bias[co];
  }
}

// dot += sum G .* W_orig   (G given in prepared layout)
__global__ void wgrad_dot_kernel(const float* __restrict__ gp, const float* __restrict__ w,
                                 float* __restrict__ dot, int Cout, int Cin, int taps, int ps_r) {
  __shared__ float scratch[32];
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    const int tap = static_cast<int>((i / Cin) % taps);
    const int cop = static_cast<int>(i / (static_cast<long long>(Cin) * taps));
    const int co = unpermute_row(cop, Cout, ps_r);
    acc = fmaf(gp[i], w[(static_cast<size_t>(co) * Cin + ci) * taps + tap], acc);
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(dot, acc);
}
// dW_orig[co,ci,tap] (+)= G/sigma - dot/sigma^2 * u[co]*v[ci*taps+tap]      (sigma == null: dW = G)
__global__ void wgrad_finish_kernel(const float* __restrict__ gp, const float* __restrict__ u,
                                    const float* __restrict__ v, const float* __restrict__ sigma,
                                    const float* __restrict__ dot, float* __restrict__ dw,
                                    const float* __restrict__ dbias_perm, float* __restrict__ dbias,
                                    int Cout, int Cin, int taps, int ps_r, int accumulate) {
  const float inv = sigma ? 1.f / *sigma : 1.f;
  const float coef = sigma ? (*dot) * inv * inv : 0.f;
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    const int tap = static_cast<int>((i / Cin) % taps);
    const int cop = static_cast<int>(i / (static_cast<long long>(Cin) * taps));
    const int co = unpermute_row(cop, Cout, ps_r);
    float g = gp[i] * inv;
    if (sigma) g -= coef * u[co] * v[ci * taps + tap];
    const size_t o = (static_cast<size_t>(co) * Cin + ci) * taps + tap;
    dw[o] = accumulate ? dw[o] + g : g;
    if (dbias && tap == 0 && ci == 0) dbias[co] = accumulate ? dbias[co] + dbias_perm[cop] : dbias_perm[cop];
  }
}

// ------------------------------------------------------------------ batched (table-driven) variants
// One launch covers every spectral-normed conv of a network.  The table lives in device memory.
__device__ __forceinline__ int find_layer(const int* __restrict__ begin, int n, int b) {
  int lo = 0, hi = n - 1;       // begin[l] <= b < begin[l+1]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (begin[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// t[k] = sum_co W[co,k] u[co] for a 32-column chunk of one layer (training layers only): 8 row groups
// per block keep the serial depth at cout/8 loads per thread
constexpr int kWtuCols = 32;
__global__ void __launch_bounds__(256)
sn_batched_wtu_kernel(const __grid_constant__ SnTable tab) {
  __shared__ float part[8][kWtuCols];
  const int* blk_begin = tab.wtu_begin;
  const int l = find_layer(blk_begin, tab.n, blockIdx.x);
  const SnLayer& L = tab.L[l];
  if (!L.training) return;
  const int col = threadIdx.x & (kWtuCols - 1), grp = threadIdx.x / kWtuCols;
  const int k = (blockIdx.x - blk_begin[l]) * kWtuCols + col;
  float acc = 0.f;
  if (k < L.K) {
    const int per = (L.cout + 7) / 8;
    const int c0 = grp * per, c1 = min(L.cout, c0 + per);
#pragma unroll 8
    for (int co = c0; co < c1; ++co) acc = fmaf(L.w[static_cast<size_t>(co) * L.K + k], L.u[co], acc);
  }
  part[grp][col] = acc;
  __syncthreads();
  if (grp == 0 && k < L.K) {
#pragma unroll
    for (int g = 1; g < 8; ++g) acc += part[g][col];
    L.t[k] = acc;
  }
}
// s_raw[co] = sum_k W[co,k] * (training ? t[k] : v[k]); one block per (layer, row)
__global__ void __launch_bounds__(128)
sn_batched_wv_kernel(const __grid_constant__ SnTable tab) {
  __shared__ float scratch[32];
  const int* row_begin = tab.row_begin;
  const int l = find_layer(row_begin, tab.n, blockIdx.x);
  const SnLayer& L = tab.L[l];
  const int co = blockIdx.x - row_begin[l];
  const float* vec = L.training ? L.t : L.v;
  const float* wr = L.w + static_cast<size_t>(co) * L.K;
  float acc = 0.f;
  for (int k = threadIdx.x; k < L.K; k += blockDim.x) acc = fmaf(wr[k], vec[k], acc);
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) L.s[co] = acc;
}
// per layer: normalise, write u/v (training), sigma, and the copies kept for backward
__global__ void __launch_bounds__(256)
sn_batched_finish_kernel(const __grid_constant__ SnTable tab, float eps) {
  __shared__ float scratch[32];
  const SnLayer& L = tab.L[blockIdx.x];
  if (L.training) {
    float a = 0.f;
    for (int k = threadIdx.x; k < L.K; k += blockDim.x) a += L.t[k] * L.t[k];
    const float nt = fmaxf(sqrtf(block_sum(a, scratch)), eps);
    for (int k = threadIdx.x; k < L.K; k += blockDim.x) {
      const float vv = L.t[k] / nt;
      L.v[k] = vv;
      L.v_saved[k] = vv;
    }
    float b = 0.f;
    for (int c = threadIdx.x; c < L.cout; c += blockDim.x) {
      const float sv = L.s[c] / nt;
      b += sv * sv;
    }
    const float ss = block_sum(b, scratch);
    const float ns = fmaxf(sqrtf(ss), eps);
    for (int c = threadIdx.x; c < L.cout; c += blockDim.x) {
      const float uu = (L.s[c] / nt) / ns;
      L.u[c] = uu;
      L.u_saved[c] = uu;
    }
    if (threadIdx.x == 0) *L.sigma = ss / ns;
  } else {
    float b = 0.f;
    for (int c = threadIdx.x; c < L.cout; c += blockDim.x) {
      b += L.u[c] * L.s[c];
      L.u_saved[c] = L.u[c];
    }
    for (int k = threadIdx.x; k < L.K; k += blockDim.x) L.v_saved[k] = L.v[k];
    b = block_sum(b, scratch);
    if (threadIdx.x == 0) *L.sigma = b;
  }
}

// weight_prep for every conv of a network.  Layers with 3x3 taps and channel counts that are multiples
// of 64 are processed as (16 co') x (64 ci) x 9 slabs through shared memory, so that the fp32 master
// rows are read, and both bf16 layouts written, in full 128-byte lines; the other (3-channel) layers
// use one thread per element.  block -> (layer, slab | chunk of 1024 elements)
constexpr int kPrepCo = 16;   // rows (co') per slab
__device__ __forceinline__ bool prep_tiled(const PrepLayer& L) {
  return L.k == 3 && L.cin % 64 == 0 && L.cout % 64 == 0;
}
__global__ void __launch_bounds__(256)
weight_prep_batched_kernel(const __grid_constant__ PrepTable tab) {
  extern __shared__ __align__(16) uint8_t prep_smem[];
  const int* blk_begin = tab.blk_begin;
  const int l = find_layer(blk_begin, tab.n, blockIdx.x);
  const PrepLayer& L = tab.L[l];
  const float inv = L.sigma ? 1.f / *L.sigma : 1.f;
  const int taps = L.k * L.k;
  const int unit = blockIdx.x - blk_begin[l];
  if (prep_tiled(L)) {
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(prep_smem);   // [kPrepCo co'][64 ci][9] (+pad)
    constexpr int kRow = 64 * 9 + 2;
    const int cib = unit % (L.cin / 64), cob = unit / (L.cin / 64);      // cob: block of kPrepCo rows
    for (int i = threadIdx.x; i < kPrepCo * 576; i += blockDim.x) {
      const int r = i / 576, j = i - r * 576;                           // j = ci_local * 9 + tap
      const int co = unpermute_row(cob * kPrepCo + r, L.cout, L.ps_r);
      tile[r * kRow + j] =
          __float2bfloat16_rn(L.w[(static_cast<size_t>(co) * L.cin + cib * 64) * 9 + j] * inv);
    }
    __syncthreads();
    // wf[co'][tap][ci]: rows of 64 ci (128 B)
    for (int i = threadIdx.x; i < kPrepCo * 576; i += blockDim.x) {
      const int ci = i & 63, tap = (i >> 6) % 9, r = i / 576;
      L.wf[(static_cast<size_t>(cob * kPrepCo + r) * 9 + tap) * L.cin + cib * 64 + ci] =
          tile[r * kRow + ci * 9 + tap];
    }
    // wd[ci][tap][co']: runs of kPrepCo co' (32 B)
    if (L.wd)
      for (int i = threadIdx.x; i < kPrepCo * 576; i += blockDim.x) {
        const int r = i % kPrepCo, tap = (i / kPrepCo) % 9, ci = i / (kPrepCo * 9);
        L.wd[(static_cast<size_t>(cib * 64 + ci) * 9 + tap) * L.cout + cob * kPrepCo + r] =
            tile[r * kRow + ci * 9 + tap];
      }
    if (L.bias_perm && cib == 0)
      for (int r = threadIdx.x; r < kPrepCo; r += blockDim.x)
        L.bias_perm[cob * kPrepCo + r] = L.bias[unpermute_row(cob * kPrepCo + r, L.cout, L.ps_r)];
    return;
  }
  const long long total = static_cast<long long>(L.cout) * taps * L.cin;
  const long long base = static_cast<long long>(unit) * 1024;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    if (i >= total) break;
    const int ci = static_cast<int>(i % L.cin);
    const int tap = static_cast<int>((i / L.cin) % taps);
    const int cop = static_cast<int>(i / (static_cast<long long>(L.cin) * taps));
    const int co = unpermute_row(cop, L.cout, L.ps_r);
    const __nv_bfloat16 h = __float2bfloat16_rn(L.w[(static_cast<size_t>(co) * L.cin + ci) * taps + tap] * inv);
    L.wf[i] = h;
    if (L.wd) L.wd[(static_cast<size_t>(ci) * taps + tap) * L.cout + cop] = h;
    if (L.bias_perm && tap == 0 && ci == 0) L.bias_perm[cop] = L.bias[co];
  }
}

// Variant for small layers (few (co, 64 ci) slabs, many splits - the generator trunk): one float4
// column of the prepared layout per thread group keeps 576 CTAs busy; dw is written with a 9-float
// stride, which is harmless on a 147 KB tensor.
// Split-K reduce + spectral-norm gradient + layout change in ONE cooperative kernel:
//   phase 1: G = sum_k partial_k (prepared layout [co'][tap][ci]); dw[co][ci][tap] (+)= G / sigma;
//            dot += <G, W_orig>; bias gradient copied
//   grid sync
//   phase 2: dw -= dot / sigma^2 * u v^T            (gradient through sigma, u and v constant)
// Block = 16 float4 columns x 16 split lanes.  phases: 3 = both (cooperative launch), 1 / 2 = one
// phase per ordinary launch.
__global__ void __launch_bounds__(256)
wgrad_reduce_finish_small_kernel(const float* __restrict__ partials, int splits, const float* __restrict__ w,
                           const float* __restrict__ u, const float* __restrict__ v,
                           const float* __restrict__ sigma, float* __restrict__ dot,
                           float* __restrict__ dw, const float* __restrict__ dbias_perm,
                           float* __restrict__ dbias, int Cout, int Cin, int taps, int ps_r,
                           int accumulate, int phases) {
  __shared__ float4 s_part[16][16];
  __shared__ float s_dot[16];
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  const long long total4 = total / 4;
  const float inv = sigma ? 1.f / *sigma : 1.f;
  if (phases & 1) {
    const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
    float dacc = 0.f;
    for (long long c0 = static_cast<long long>(blockIdx.x) * 16; c0 < total4;
         c0 += static_cast<long long>(gridDim.x) * 16) {
      const long long i4 = c0 + x;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i4 < total4) {
        const float4* p = reinterpret_cast<const float4*>(partials) + i4;
#pragma unroll 4
        for (int k = y; k < splits; k += 16) {
          const float4 t = __ldg(p + static_cast<long long>(k) * total4);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
      }
      s_part[y][x] = acc;
      __syncthreads();
      if (y == 0 && i4 < total4) {
#pragma unroll
        for (int k = 1; k < 16; ++k) {
          const float4 t = s_part[k][x];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const long long i = i4 * 4;
        const int ci = static_cast<int>(i % Cin);
        const int tap = static_cast<int>((i / Cin) % taps);
        const int cop = static_cast<int>(i / (static_cast<long long>(Cin) * taps));
        const int co = unpermute_row(cop, Cout, ps_r);
        const size_t o = (static_cast<size_t>(co) * Cin + ci) * taps + tap;
        const float g[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const size_t oe = o + static_cast<size_t>(e) * taps;
          if (sigma) dacc = fmaf(g[e], w[oe], dacc);
          dw[oe] = accumulate ? dw[oe] + g[e] * inv : g[e] * inv;
        }
        if (dbias && tap == 0 && ci == 0)
          dbias[co] = accumulate ? dbias[co] + dbias_perm[cop] : dbias_perm[cop];
      }
      __syncthreads();
    }
    if (sigma) {
      if (y == 0) s_dot[x] = dacc;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 16; ++k) t += s_dot[k];
        atomicAdd(dot, t);
      }
    }
  }
  if (phases == 3) cooperative_groups::this_grid().sync();
  if ((phases & 2) && sigma) {
    const float coef = (*reinterpret_cast<volatile float*>(dot)) * inv * inv;
    const int K = Cin * taps;
    for (long long o = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; o < total;
         o += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int co = static_cast<int>(o / K);
      const int k = static_cast<int>(o - static_cast<long long>(co) * K);
      dw[o] -= coef * u[co] * v[k];
    }
  }
}

// Split-K reduce + spectral-norm gradient + layout change in ONE cooperative kernel:
//   phase 1: G = sum_k partial_k (prepared layout [co'][tap][ci]); dw[co][ci][tap] (+)= G / sigma;
//            dot += <G, W_orig>; bias gradient copied
//   grid sync
//   phase 2: dw -= dot / sigma^2 * u v^T            (gradient through sigma, u and v constant)
// Block = 16 float4 columns x 16 split lanes.  phases: 3 = both (cooperative launch), 1 / 2 = one
// phase per ordinary launch.
__global__ void __launch_bounds__(256)
wgrad_reduce_finish_kernel(const float* __restrict__ partials, int splits, const float* __restrict__ w,
                           const float* __restrict__ u, const float* __restrict__ v,
                           const float* __restrict__ sigma, float* __restrict__ dot,
                           float* __restrict__ dw, const float* __restrict__ dbias_perm,
                           float* __restrict__ dbias, int Cout, int Cin, int taps, int ps_r,
                           int accumulate, int phases) {
  __shared__ float s_tile[64 * 9];      // one (co, 64 ci) slab in master order [ci][tap]
  __shared__ float s_dot[8];
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  const long long total4 = total / 4;
  const float inv = sigma ? 1.f / *sigma : 1.f;
  if (phases & 1) {
    // work item = (prepared row co', 64-channel chunk): its 9 x 64 values are contiguous in BOTH layouts
    // ([tap][ci] in the partials, [ci][tap] in dw / w), so the layout change is a shared-memory transpose
    // and every global access is coalesced.
    // (used for the large layers, which have few splits: one thread per float4 column of the slab)
    const int chunks = Cin / 64;
    const int tap_t = threadIdx.x >> 4, x = threadIdx.x & 15;
    float dacc = 0.f;
    for (int item = blockIdx.x; item < Cout * chunks; item += gridDim.x) {
      const int cop = item / chunks, cic = item - cop * chunks;
      const int co = unpermute_row(cop, Cout, ps_r);
      if (tap_t < taps) {
        const long long i4 = ((static_cast<long long>(cop) * taps + tap_t) * Cin + cic * 64) / 4 + x;
        const float4* p = reinterpret_cast<const float4*>(partials) + i4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < splits; ++k) {
          const float4 t = __ldg(p + static_cast<long long>(k) * total4);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        s_tile[(4 * x + 0) * taps + tap_t] = acc.x;
        s_tile[(4 * x + 1) * taps + tap_t] = acc.y;
        s_tile[(4 * x + 2) * taps + tap_t] = acc.z;
        s_tile[(4 * x + 3) * taps + tap_t] = acc.w;
      }
      __syncthreads();
      const size_t base = (static_cast<size_t>(co) * Cin + cic * 64) * taps;
      for (int i = threadIdx.x; i < 64 * taps; i += blockDim.x) {
        const float g = s_tile[i];
        if (sigma) dacc = fmaf(g, w[base + i], dacc);
        dw[base + i] = accumulate ? dw[base + i] + g * inv : g * inv;
      }
      if (dbias && cic == 0 && threadIdx.x == 0)
        dbias[co] = accumulate ? dbias[co] + dbias_perm[cop] : dbias_perm[cop];
      __syncthreads();
    }
    if (sigma) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
      if ((threadIdx.x & 31) == 0) s_dot[threadIdx.x >> 5] = dacc;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += s_dot[k];
        atomicAdd(dot, t);
      }
    }
  }
  if (phases == 3) cooperative_groups::this_grid().sync();
  if ((phases & 2) && sigma) {
    const float coef = (*reinterpret_cast<volatile float*>(dot)) * inv * inv;
    const int K = Cin * taps;
    for (long long o = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; o < total;
         o += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int co = static_cast<int>(o / K);
      const int k = static_cast<int>(o - static_cast<long long>(co) * K);
      dw[o] -= coef * u[co] * v[k];
    }
  }
}

inline int grid_for(long long work) {
  long long b = (work + kThreads - 1) / kThreads;
  if (b > 148 * 8) b = 148 * 8;
  return static_cast<int>(b < 1 ? 1 : b);
}
int check() { return cudaGetLastError() == cudaSuccess ? 0 : 4; }
bool g_no_coop = false;

}  // namespace

void weight_grad_disable_cooperative(int off) { g_no_coop = off != 0; }

size_t sn_workspace_floats(int Cout, int K) { return static_cast<size_t>(K) + Cout + 4; }

int sn_power_iteration(const float* w, float* u, float* v, float* sigma, int Cout, int K,
                       int training, float eps, float* ws, cudaStream_t s) {
  float* t = ws;
  float* s_raw = ws + K;
  if (training) {
    cudaMemsetAsync(t, 0, sizeof(float) * K, s);
    const int co_chunk = 64;
    dim3 grid((K + kThreads - 1) / kThreads, (Cout + co_chunk - 1) / co_chunk);
    sn_wtu_kernel<<<grid, kThreads, 0, s>>>(w, u, t, Cout, K, co_chunk);
    sn_wv_kernel<<<Cout, kThreads, 0, s>>>(w, t, s_raw, K);
  } else {
    sn_wv_kernel<<<Cout, kThreads, 0, s>>>(w, v, s_raw, K);
  }
  sn_finish_kernel<<<1, kThreads, 0, s>>>(t, s_raw, u, v, sigma, Cout, K, training, eps);
  return check();
}

int sn_power_iteration_batched(const SnLayer* layers, int n_layers, float eps, cudaStream_t s) {
  for (int base = 0; base < n_layers; base += kMaxBatch) {
    SnTable tab;
    tab.n = n_layers - base < kMaxBatch ? n_layers - base : kMaxBatch;
    int wtu = 0, rows = 0;
    for (int i = 0; i < tab.n; ++i) {
      tab.L[i] = layers[base + i];
      tab.wtu_begin[i] = wtu;
      tab.row_begin[i] = rows;
      if (tab.L[i].training) wtu += (tab.L[i].K + kWtuCols - 1) / kWtuCols;
      rows += tab.L[i].cout;
    }
    tab.wtu_begin[tab.n] = wtu;
    tab.row_begin[tab.n] = rows;
    if (wtu > 0) sn_batched_wtu_kernel<<<wtu, 256, 0, s>>>(tab);
    sn_batched_wv_kernel<<<rows, 128, 0, s>>>(tab);
    sn_batched_finish_kernel<<<tab.n, 256, 0, s>>>(tab, eps);
  }
  return check();
}

int weight_prep_batched(const PrepLayer* layers, int n_layers, cudaStream_t s) {
  for (int base = 0; base < n_layers; base += kMaxBatch) {
    PrepTable tab;
    tab.n = n_layers - base < kMaxBatch ? n_layers - base : kMaxBatch;
    int blocks = 0;
    for (int i = 0; i < tab.n; ++i) {
      tab.L[i] = layers[base + i];
      if (tab.L[i].ps_r > 1 && tab.L[i].cout % (tab.L[i].ps_r * tab.L[i].ps_r)) return 1;
      tab.blk_begin[i] = blocks;
      const PrepLayer& Li = tab.L[i];
      const long long total = static_cast<long long>(Li.cout) * Li.cin * Li.k * Li.k;
      if (Li.k == 3 && Li.cin % 64 == 0 && Li.cout % 64 == 0)
        blocks += (Li.cout / kPrepCo) * (Li.cin / 64);
      else
        blocks += static_cast<int>((total + 1023) / 1024);
    }
    tab.blk_begin[tab.n] = blocks;
    constexpr int kPrepSmem = kPrepCo * (64 * 9 + 2) * 2;
    static bool configured = false;
    if (!configured) {
      if (cudaFuncSetAttribute(weight_prep_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kPrepSmem) != cudaSuccess)
        return 3;
      configured = true;
    }
    if (blocks > 0) weight_prep_batched_kernel<<<blocks, 256, kPrepSmem, s>>>(tab);
  }
  return check();
}

int weight_prep(const float* w, const float* sigma, const float* bias, __nv_bfloat16* wf,
                __nv_bfloat16* wd, float* bias_perm, int Cout, int Cin, int KH, int KW, int ps_r,
                cudaStream_t s) {
  if (ps_r > 1 && Cout % (ps_r * ps_r)) return 1;
  const long long total = static_cast<long long>(Cout) * Cin * KH * KW;
  weight_prep_kernel<<<grid_for(total), kThreads, 0, s>>>(w, sigma, bias, wf, wd, bias_perm, Cout, Cin,
                                                          KH, KW, ps_r);
  return check();
}

int weight_grad_reduce_finish(const float* partials, int splits, const float* w, const float* u,
                              const float* v, const float* sigma, float* dw, const float* dbias_perm,
                              float* dbias, int Cout, int Cin, int KH, int KW, int ps_r, int accumulate,
                              float* dot, cudaStream_t s) {
  int taps = KH * KW;
  if (Cin % 64 || taps > 9) return 1;
  if (sigma) cudaMemsetAsync(dot, 0, sizeof(float), s);
  static int max_coop_blocks = -1;   // co-resident blocks of the cooperative kernel (0: unsupported)
  if (max_coop_blocks < 0) {
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wgrad_reduce_finish_kernel, 256, 0);
    int per_sm2 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, wgrad_reduce_finish_small_kernel, 256, 0);
    if (per_sm2 < per_sm) per_sm = per_sm2;
    max_coop_blocks = (coop && per_sm > 0 && sms > 0) ? sms * (per_sm < 4 ? per_sm : 4) : 0;
    cudaGetLastError();
  }
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  const long long slabs = static_cast<long long>(Cout) * (Cin / 64);
  const bool small = slabs < 512;      // trunk-sized layers: parallelism over coalescing
  auto kernel = small ? wgrad_reduce_finish_small_kernel : wgrad_reduce_finish_kernel;
  long long want = small ? (total / 4 + 15) / 16 : slabs;
  if (!sigma) {   // no gradient through sigma: phase 1 alone is complete
    const int grid = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
    kernel<<<grid, 256, 0, s>>>(partials, splits, w, u, v, sigma, dot, dw, dbias_perm, dbias, Cout, Cin, taps,
                                ps_r, accumulate, 1);
    return check();
  }
  if (max_coop_blocks > 0 && !g_no_coop) {
    const int grid = static_cast<int>(want < max_coop_blocks ? want : max_coop_blocks);
    int phases = 3;
    void* args[] = {&partials, &splits, &w, &u, &v, &sigma, &dot, &dw, &dbias_perm, &dbias,
                    &Cout, &Cin, &taps, &ps_r, &accumulate, &phases};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel),
                                                dim3(grid), dim3(256), args, 0, s);
    return e == cudaSuccess ? 0 : 4;
  }
  const int grid = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
  kernel<<<grid, 256, 0, s>>>(partials, splits, w, u, v, sigma, dot, dw, dbias_perm, dbias, Cout, Cin, taps, ps_r,
                              accumulate, 1);
  kernel<<<grid, 256, 0, s>>>(partials, splits, w, u, v, sigma, dot, dw, dbias_perm, dbias, Cout, Cin, taps, ps_r,
                              accumulate, 2);
  return check();
}

int weight_grad_finish(const float* gp, const float* w, const float* u, const float* v,
                       const float* sigma, float* dw, const float* dbias_perm, float* dbias, int Cout,
                       int Cin, int KH, int KW, int ps_r, int accumulate, float* ws, cudaStream_t s) {
  const long long total = static_cast<long long>(Cout) * Cin * KH * KW;
  float* dot = ws;
  if (sigma) {
    cudaMemsetAsync(dot, 0, sizeof(float), s);
    wgrad_dot_kernel<<<grid_for(total), kThreads, 0, s>>>(gp, w, dot, Cout, Cin, KH * KW, ps_r);
  }
  wgrad_finish_kernel<<<grid_for(total), kThreads, 0, s>>>(gp, u, v, sigma, dot, dw, dbias_perm, dbias,
                                                           Cout, Cin, KH * KW, ps_r, accumulate);
  return check();
}

}  // namespace sisr
