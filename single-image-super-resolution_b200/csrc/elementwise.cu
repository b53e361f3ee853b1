// HBM/L2-bound kernels of the SRGAN step on NHWC bf16 activations:
// layout conversion at the module boundary, BatchNorm statistics / apply / backward fused with
// PReLU / LeakyReLU / residual add, activation backward for epilogue-fused activations,
// 2x2 max-pool, feature-MSE and BCE losses.  16-byte vector accesses, fp32 arithmetic.
#include "elementwise.h"

#include <stdio.h>

#include "peer_device.cuh"
#include "ptx.cuh"

namespace sisr {

namespace {

constexpr int kThreads = 256;

struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Vec8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ Vec8 cvt8(const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Vec8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
constexpr int kRowsPerTrip = 4;   // independent 16-byte load pairs in flight per thread in the BN backward loops
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]);
  u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]);
  u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  return (act == ACT_NONE || z > 0.f) ? z : z * slope;
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
  return (act == ACT_NONE || z > 0.f) ? 1.f : slope;
}
__device__ __forceinline__ float resolve_slope(int act, float slope, const float* slope_ptr) {
  if (act == ACT_PRELU) return *slope_ptr;
  if (act == ACT_RELU) return 0.f;
  return slope;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline int grid_for(long long work, int threads, int cap = 148 * 16) {
  long long b = (work + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------ layout conversion
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                             int N, int C, int H, int W) {
  const long long total = static_cast<long long>(N) * C * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long p = i / C;
    const int w = static_cast<int>(p % W);
    p /= W;
    const int h = static_cast<int>(p % H);
    const int n = static_cast<int>(p / H);
    y[i] = __float2bfloat16_rn(x[((static_cast<long long>(n) * C + c) * H + h) * W + w]);
  }
}
// images (C <= 4): one thread per pixel, 64-bit pixel indices avoided (one division per thread)
template <int C>
__global__ void nchw_f32_to_nhwc_bf16_small_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                   long long npix, int HW) {
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < npix;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = q / HW;
    const float* src = x + (n * C) * HW + (q - n * HW);
#pragma unroll
    for (int c = 0; c < C; ++c) y[q * C + c] = __float2bfloat16_rn(src[static_cast<size_t>(c) * HW]);
  }
}
template <int C>
__global__ void nhwc_bf16_to_nchw_f32_small_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                                   long long npix, int HW) {
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < npix;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = q / HW;
    float* dst = y + (n * C) * HW + (q - n * HW);
#pragma unroll
    for (int c = 0; c < C; ++c) dst[static_cast<size_t>(c) * HW] = __bfloat162float(x[q * C + c]);
  }
}
// tiled transpose per image: [HW, C] bf16 -> [C, HW] f32 (and back)
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                             int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * HW * C;
  float* yn = y + static_cast<size_t>(n) * HW * C;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[r][threadIdx.x] = __bfloat162float(xn[static_cast<size_t>(p) * C + c]);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) yn[static_cast<size_t>(c) * HW + p] = tile[threadIdx.x][r];
  }
}
__global__ void nchw_f32_to_nhwc_bf16_tiled_kernel(const float* __restrict__ x,
                                                   __nv_bfloat16* __restrict__ y, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xn = x + static_cast<size_t>(n) * HW * C;
  __nv_bfloat16* yn = y + static_cast<size_t>(n) * HW * C;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) tile[r][threadIdx.x] = xn[static_cast<size_t>(c) * HW + p];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < C) yn[static_cast<size_t>(p) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][r]);
  }
}
// [HW, C] bf16 <-> [C, HW] bf16 (discriminator flatten order)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                      int R, int Cc) {
  // x: [n][R][Cc] -> y: [n][Cc][R]
  __shared__ __nv_bfloat16 tile[32][34];
  const int n = blockIdx.z;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * R * Cc;
  __nv_bfloat16* yn = y + static_cast<size_t>(n) * R * Cc;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = xn[static_cast<size_t>(r) * Cc + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) yn[static_cast<size_t>(c) * R + r] = tile[threadIdx.x][i];
  }
}

// dpre[n,h,w,c] (bf16) = dout[n,c,h,w] * (1 - y[n,c,h,w]^2)
__global__ void tanh_bwd_nchw_to_nhwc_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                             __nv_bfloat16* __restrict__ dpre, int N, int C, int H,
                                             int W) {
  const long long total = static_cast<long long>(N) * C * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long p = i / C;
    const int w = static_cast<int>(p % W);
    p /= W;
    const int h = static_cast<int>(p % H);
    const int n = static_cast<int>(p / H);
    const long long j = ((static_cast<long long>(n) * C + c) * H + h) * W + w;
    const float yv = y[j];
    dpre[i] = __float2bfloat16_rn(dout[j] * (1.f - yv * yv));
  }
}

// ------------------------------------------------------------------ per-channel reductions
// stats[0..C) += sum_rows y, stats[C..2C) += sum_rows y^2  (y: [M, C] bf16, C % 8 == 0, C <= 2048)
__global__ void col_stats_kernel(const __nv_bfloat16* __restrict__ y, long long M, int C,
                                 float* __restrict__ stats, int with_sq) {
  extern __shared__ float s_acc[];  // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int tpr = C / 8;
  const int rpb = blockDim.x / tpr;
  const int r = threadIdx.x / tpr, cv = threadIdx.x % tpr;
  float s1[8] = {0}, s2[8] = {0};
  if (r < rpb) {
    for (long long row = static_cast<long long>(blockIdx.x) * rpb + r; row < M;
         row += static_cast<long long>(gridDim.x) * rpb) {
      const Vec8 v = load8(y + row * C + cv * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += v.v[j];
        s2[j] += v.v[j] * v.v[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_acc[cv * 8 + j], s1[j]);
      if (with_sq) atomicAdd(&s_acc[C + cv * 8 + j], s2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (with_sq ? 2 * C : C); i += blockDim.x) atomicAdd(&stats[i], s_acc[i]);
}

// ------------------------------------------------------------------ BatchNorm finalize
// One block per 8 channels: 32 row lanes x 8 channels add the per-CTA partial rows written by the
// conv epilogue (independent loads), shared-memory tree over the row lanes, 8 threads finalize.
__global__ void __launch_bounds__(kThreads)
bn_finalize_kernel(const float* __restrict__ stats, int stats_rows, float count,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ num_batches, float momentum, float eps, int training,
                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                   float* __restrict__ invstd_out, int C) {
  __shared__ float s1[32][8], s2[32][8];
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && num_batches) *num_batches += 1;
  float a1 = 0.f, a2 = 0.f;
  if (training && c < C) {
#pragma unroll 5
    for (int r = rl; r < stats_rows; r += 32) {
      a1 += stats[static_cast<size_t>(r) * 2 * C + c];
      a2 += stats[static_cast<size_t>(r) * 2 * C + C + c];
    }
  }
  s1[rl][cl] = a1;
  s2[rl][cl] = a2;
  __syncthreads();
  if (rl != 0 || c >= C) return;
  float mean, var;
  if (training) {
    for (int r = 1; r < 32; ++r) {
      a1 += s1[r][cl];
      a2 += s2[r][cl];
    }
    mean = a1 / count;
    var = fmaxf(a2 / count - mean * mean, 0.f);
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
}

// ------------------------------------------------------------------ BatchNorm apply (+act, +residual)
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, int act, float slope,
                                const float* __restrict__ slope_ptr,
                                const __nv_bfloat16* __restrict__ residual,
                                __nv_bfloat16* __restrict__ out, long long nvec, int C) {
  pdl_launch_dependents();
  constexpr int U = 4;       // independent 16-byte loads (pairs with a residual) in flight per thread
  const float sl = resolve_slope(act, slope, slope_ptr);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < nvec;
       i0 += U * stride) {
    uint4 vv[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        vv[u] = *reinterpret_cast<const uint4*>(y + i * 8);
        if (residual) rr[u] = *reinterpret_cast<const uint4*>(residual + i * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) break;
      const int c0 = static_cast<int>((i * 8) % C);
      const float4 sc0 = *reinterpret_cast<const float4*>(scale + c0);
      const float4 sc1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
      const float4 sh0 = *reinterpret_cast<const float4*>(shift + c0);
      const float4 sh1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
      const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
      const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
      Vec8 v = cvt8(vv[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = act_fwd(fmaf(v.v[j], sc[j], sh[j]), act, sl);
      if (residual) {
        const Vec8 r = cvt8(rr[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] += r.v[j];
      }
      store8(out + i * 8, v);
    }
  }
}

// ------------------------------------------------------------------ per-channel block reductions
// Thread (r, cv) of a 256-thread block owns channels [8cv, 8cv+8) of the rows r, r+rpb, ... it
// visits (tpr = C/8 threads per row, rpb = 256/tpr rows in flight).  After the row loop the NQ
// per-thread partial vectors are combined without atomics: staged as s_red[q][r][C], summed over r
// by one thread per column, and added to global memory once per column per block.
constexpr int kRedFloats = 2048;   // rpb * C for every supported C (C % 8 == 0, C <= 2048)
// (A variant that combined 8 CTAs per thread-block cluster through distributed shared memory before
// touching global memory was measured 2x SLOWER on the 4.7 MB trunk tensors - cluster scheduling and
// cluster.sync cost more than the atomics they save - profiles/r1_notes.md.)
template <int NQ>
__device__ __forceinline__ void block_col_reduce(const float (&p)[NQ][8], int r, int cv, int rpb, int C,
                                                 bool active, float* s_red, float* const (&out)[NQ]) {
  if (active) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float4* d = reinterpret_cast<float4*>(s_red + (q * rpb + r) * C + cv * 8);
      d[0] = make_float4(p[q][0], p[q][1], p[q][2], p[q][3]);
      d[1] = make_float4(p[q][4], p[q][5], p[q][6], p[q][7]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * C; i += blockDim.x) {
    const int q = i / C, c = i - q * C;
    if (!out[q]) continue;
    float acc = 0.f;
    for (int rr = 0; rr < rpb; ++rr) acc += s_red[(q * rpb + rr) * C + c];
    atomicAdd(out[q] + c, acc);
  }
}

// ------------------------------------------------------------------ BatchNorm backward
// sums[0..C) += sum g, sums[C..2C) += sum g*xhat, sums[2C] += sum dout*min(0,z)   (g = dout*act'(z))
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     const float* __restrict__ scale, const float* __restrict__ shift, int act,
                     float slope, const float* __restrict__ slope_ptr, float* __restrict__ sums,
                     long long M, int C, const __grid_constant__ PeerTable peer, int slot,
                     float* __restrict__ sums_global, unsigned int* __restrict__ ticket) {
  __shared__ __align__(16) float s_red[2 * kRedFloats];
  __shared__ float s_sa[kThreads / 32];
  const float sl = resolve_slope(act, slope, slope_ptr);
  const int tpr = C / 8;
  const int rpb = blockDim.x / tpr;
  const int r = threadIdx.x / tpr, cv = threadIdx.x % tpr;
  const bool active = r < rpb;
  float p[2][8] = {};
  float sa = 0.f;
  if (active) {
    float mu[8], is[8], sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = mean[cv * 8 + j];
      is[j] = invstd[cv * 8 + j];
      sc[j] = scale[cv * 8 + j];
      sh[j] = shift[cv * 8 + j];
    }
    // kRowsPerTrip rows per trip, kept as raw 16-byte words until used
    constexpr int U = kRowsPerTrip;
    const long long stride = static_cast<long long>(gridDim.x) * rpb;
    for (long long row0 = static_cast<long long>(blockIdx.x) * rpb + r; row0 < M; row0 += U * stride) {
      uint4 dr[U], vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long row = row0 + u * stride;
        if (row < M) {
          dr[u] = *reinterpret_cast<const uint4*>(dout + row * C + cv * 8);
          vr[u] = *reinterpret_cast<const uint4*>(y + row * C + cv * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (row0 + u * stride >= M) break;
        const Vec8 d = cvt8(dr[u]), v = cvt8(vr[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(v.v[j], sc[j], sh[j]);
          const float g = d.v[j] * act_grad(z, act, sl);
          p[0][j] += g;
          p[1][j] += g * (v.v[j] - mu[j]) * is[j];
          if (act == ACT_PRELU) sa += d.v[j] * fminf(z, 0.f);
        }
      }
    }
  }
  float* const outs[2] = {sums, sums + C};
  block_col_reduce<2>(p, r, cv, rpb, C, active, s_red, outs);
  if (act == ACT_PRELU) {
    sa = warp_sum(sa);
    if ((threadIdx.x & 31) == 0) s_sa[threadIdx.x >> 5] = sa;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < kThreads / 32; ++i) t += s_sa[i];
      atomicAdd(&sums[2 * C], t);
    }
  }
  // SyncBN (world > 1): the last CTA to finish exchanges the [2C+1] local sums with the peers over
  // NVLink and writes the global sums (no separate all-reduce launch)
  if (peer.world > 1) {
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      float* vals = s_red;                          // 2C+1 <= 2 * kRedFloats
      const int n = 2 * C + 1;
      for (int i = threadIdx.x; i < n; i += blockDim.x) vals[i] = __ldcg(sums + i);
      __syncthreads();
      exchange(peer, slot, vals, n);
      for (int i = threadIdx.x; i < n; i += blockDim.x) sums_global[i] = vals[i];
    }
  }
}


// dy = gamma*invstd * (g - sum_g/count - xhat * sum_gx/count);  colsum[c] += sum_rows dy (as stored)
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ scale, const float* __restrict__ shift, int act,
                    float slope, const float* __restrict__ slope_ptr, const float* __restrict__ sums,
                    float inv_count, __nv_bfloat16* __restrict__ dy, float* __restrict__ colsum,
                    long long M, int C) {
  pdl_launch_dependents();
  __shared__ __align__(16) float s_red[kRedFloats];
  const float sl = resolve_slope(act, slope, slope_ptr);
  const int tpr = C / 8;
  const int rpb = blockDim.x / tpr;
  const int r = threadIdx.x / tpr, cv = threadIdx.x % tpr;
  const bool active = r < rpb;
  float p[1][8] = {};
  if (active) {
    float mu[8], is[8], sc[8], sh[8], k1[8], k2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      mu[j] = mean[c];
      is[j] = invstd[c];
      sc[j] = scale[c];
      sh[j] = shift[c];
      k1[j] = sums[c] * inv_count;
      k2[j] = sums[C + c] * inv_count;
    }
    const long long stride = static_cast<long long>(gridDim.x) * rpb;
    constexpr int U = kRowsPerTrip;
    for (long long row0 = static_cast<long long>(blockIdx.x) * rpb + r; row0 < M; row0 += U * stride) {
      uint4 dr[U], vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long row = row0 + u * stride;
        if (row < M) {
          dr[u] = *reinterpret_cast<const uint4*>(dout + row * C + cv * 8);
          vr[u] = *reinterpret_cast<const uint4*>(y + row * C + cv * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long row = row0 + u * stride;
        if (row >= M) break;
        const Vec8 d = cvt8(dr[u]), v = cvt8(vr[u]);
        Vec8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(v.v[j], sc[j], sh[j]);
          const float g = d.v[j] * act_grad(z, act, sl);
          const float xh = (v.v[j] - mu[j]) * is[j];
          o.v[j] = bf16_round(sc[j] * (g - k1[j] - xh * k2[j]));
          p[0][j] += o.v[j];
        }
        store8(dy + row * C + cv * 8, o);
      }
    }
  }
  if (colsum) {
    float* const outs[1] = {colsum};
    block_col_reduce<1>(p, r, cv, rpb, C, active, s_red, outs);
  }
}

// ------------------------------------------------------------------ activation backward (from output)
// din = dout * (out > 0 ? 1 : slope); dslope += sum_{out<0} dout*out/slope   (requires slope > 0);
// colsum[c] += sum_rows din (as stored)
__global__ void __launch_bounds__(kThreads)
act_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out, int act,
               float slope, const float* __restrict__ slope_ptr, __nv_bfloat16* __restrict__ din,
               float* __restrict__ dslope, float* __restrict__ colsum, long long M, int C) {
  pdl_launch_dependents();
  __shared__ __align__(16) float s_red[kRedFloats];
  __shared__ float s_part[kThreads / 32];
  const float sl = resolve_slope(act, slope, slope_ptr);
  const float inv_sl = sl != 0.f ? 1.f / sl : 0.f;
  const int tpr = C / 8;
  const int rpb = blockDim.x / tpr;
  const int r = threadIdx.x / tpr, cv = threadIdx.x % tpr;
  const bool active = r < rpb;
  float p[1][8] = {};
  float sa = 0.f;
  if (active) {
    for (long long row = static_cast<long long>(blockIdx.x) * rpb + r; row < M;
         row += static_cast<long long>(gridDim.x) * rpb) {
      const Vec8 d = load8(dout + row * C + cv * 8);
      const Vec8 o = load8(out + row * C + cv * 8);
      Vec8 q;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool pos = o.v[j] > 0.f;
        q.v[j] = bf16_round(pos ? d.v[j] : d.v[j] * sl);
        p[0][j] += q.v[j];
        if (!pos) sa += d.v[j] * o.v[j] * inv_sl;
      }
      store8(din + row * C + cv * 8, q);
    }
  }
  if (colsum) {
    float* const outs[1] = {colsum};
    block_col_reduce<1>(p, r, cv, rpb, C, active, s_red, outs);
  }
  if (dslope) {
    sa = warp_sum(sa);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sa;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < kThreads / 32 ? s_part[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) atomicAdd(dslope, t);
    }
  }
}

// ------------------------------------------------------------------ 2x2 max-pool
__global__ void maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                    int N, int H, int W, int C) {
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2, cv = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    long long p = i / cv;
    const int ow = static_cast<int>(p % OW);
    p /= OW;
    const int oh = static_cast<int>(p % OH);
    const int n = static_cast<int>(p / OH);
    const __nv_bfloat16* b = x + ((static_cast<size_t>(n) * H + oh * 2) * W + ow * 2) * C + c * 8;
    const Vec8 a0 = load8(b), a1 = load8(b + C), a2 = load8(b + static_cast<size_t>(W) * C),
               a3 = load8(b + static_cast<size_t>(W) * C + C);
    Vec8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.v[j] = fmaxf(fmaxf(a0.v[j], a1.v[j]), fmaxf(a2.v[j], a3.v[j]));
    store8(y + i * 8, r);
  }
}
// gradient goes to the first maximum in scan order (torch semantics); dx fully written for even H, W
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                    const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                    int N, int H, int W, int C) {
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2, cv = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    long long p = i / cv;
    const int ow = static_cast<int>(p % OW);
    p /= OW;
    const int oh = static_cast<int>(p % OH);
    const int n = static_cast<int>(p / OH);
    const size_t base = ((static_cast<size_t>(n) * H + oh * 2) * W + ow * 2) * C + c * 8;
    const size_t offs[4] = {0, static_cast<size_t>(C), static_cast<size_t>(W) * C,
                            static_cast<size_t>(W) * C + C};
    Vec8 a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = load8(x + base + offs[k]);
    const Vec8 g = load8(dy + i * 8);
    Vec8 o[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float m = a[0].v[j];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (a[k].v[j] > m) {
          m = a[k].v[j];
          best = k;
        }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k].v[j] = (k == best) ? g.v[j] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) store8(dx + base + offs[k], o[k]);
  }
}

// ------------------------------------------------------------------ stand-alone PixelShuffle(2)
// NHWC bf16: x [N, H, W, 4*C] -> y [N, 2H, 2W, C] with y[n, 2h+i, 2w+j, c] = x[n, h, w, c*4 + i*2 + j]
// (nn.PixelShuffle channel order).  inverse = 1: the same mapping read backwards (gradient).  Used by
// the narrow-channel stages of model_generator_progressive.py; the main generator folds the shuffle
// into the producing conv's store addressing instead.
__global__ void pixel_shuffle2_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      int N, int H, int W, int C, int inverse) {
  const long long total = static_cast<long long>(N) * H * W * 4 * C;
  for (long long o = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<long long>(gridDim.x) * blockDim.x) {
    // o indexes the shuffled tensor [N, 2H, 2W, C]
    const int c = static_cast<int>(o % C);
    long long q = o / C;
    const int ox = static_cast<int>(q % (2 * W));
    q /= 2 * W;
    const int oy = static_cast<int>(q % (2 * H));
    const int n = static_cast<int>(q / (2 * H));
    const long long i = ((static_cast<long long>(n) * H + (oy >> 1)) * W + (ox >> 1)) * 4 * C + c * 4 +
                        (oy & 1) * 2 + (ox & 1);
    if (inverse) dst[i] = src[o]; else dst[o] = src[i];
  }
}

// ------------------------------------------------------------------ LR synthesis (utils.py:16-31)
// F.interpolate(mode='bicubic', align_corners=True) (cubic convolution, A = -0.75, border-clamped
// taps) followed by the clamp to [-1, 1]; NCHW fp32 in and out.
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__device__ __forceinline__ void cubic_source(int dst, float scale, int in_size, int& idx, float& t) {
  const float real = scale * static_cast<float>(dst);
  int i = static_cast<int>(floorf(real));
  if (i > in_size - 1) i = in_size - 1;
  float lam = real - static_cast<float>(i);
  lam = fminf(fmaxf(lam, 0.f), 1.f);
  idx = i;
  t = lam;
}
// mode 0: lr = clamp(interp(hr));  mode 1 (backward): dhr += W^T (dlr masked where the interpolated
// value left [-1, 1])
__global__ void lr_from_hr_kernel(const float* __restrict__ hr, float* __restrict__ lr,
                                  const float* __restrict__ dlr, float* __restrict__ dhr, int NC, int H,
                                  int W, int OH, int OW, float sh, float sw) {
  const long long total = static_cast<long long>(NC) * OH * OW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % OW);
    const int oy = static_cast<int>((i / OW) % OH);
    const long long nc = i / (static_cast<long long>(OW) * OH);
    int ix, iy;
    float tx, ty, wx[4], wy[4];
    cubic_source(ox, sw, W, ix, tx);
    cubic_source(oy, sh, H, iy, ty);
    cubic_coeffs(tx, wx);
    cubic_coeffs(ty, wy);
    const float* src = hr + nc * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), H - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), W - 1);
        row += src[yy * W + xx] * wx[b];
      }
      acc += row * wy[a];
    }
    if (!dlr) {
      lr[i] = fminf(fmaxf(acc, -1.f), 1.f);
    } else {
      const float g = (acc > -1.f && acc < 1.f) ? dlr[i] : 0.f;
      if (g != 0.f) {
        float* dst = dhr + nc * H * W;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int yy = min(max(iy - 1 + a, 0), H - 1);
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int xx = min(max(ix - 1 + b, 0), W - 1);
            atomicAdd(dst + yy * W + xx, g * wx[b] * wy[a]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ losses
// loss += weight/n * sum (a-b)^2 ;  (fp32 inputs)
__global__ void mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                               float coef, float* __restrict__ loss) {
  float s = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = a[i] - b[i];
    s += d * d;
  }
  __shared__ float s_part[kThreads / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kThreads / 32 ? s_part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(loss, t * coef);
  }
}
// grad_b = gout * 2*coef * (b - a),  grad_a = -grad_b
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                               float coef, const float* __restrict__ gout, float* __restrict__ ga,
                               float* __restrict__ gb) {
  const float k = 2.f * coef * (*gout);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = (b[i] - a[i]) * k;
    if (gb) gb[i] = d;
    if (ga) ga[i] = -d;
  }
}
// BCELoss(mean) against a constant target with torch's clamping; dp = gout * (p-t)/max(p(1-p),1e-12)/n
__global__ void bce_kernel(const float* __restrict__ p, int n, float target, float* __restrict__ loss,
                           float* __restrict__ mean_p) {
  float s = 0.f, sp = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float q = p[i];
    s -= target * fmaxf(logf(q), -100.f) + (1.f - target) * fmaxf(logf(1.f - q), -100.f);
    sp += q;
  }
  __shared__ float s_part[2][kThreads / 32];
  s = warp_sum(s);
  sp = warp_sum(sp);
  if ((threadIdx.x & 31) == 0) {
    s_part[0][threadIdx.x >> 5] = s;
    s_part[1][threadIdx.x >> 5] = sp;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, tp = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) {
      t += s_part[0][i];
      tp += s_part[1][i];
    }
    *loss = t / n;
    if (mean_p) *mean_p = tp / n;
  }
}
__global__ void bce_bwd_kernel(const float* __restrict__ p, int n, float target,
                               const float* __restrict__ gout, float* __restrict__ dp) {
  const float g = *gout / n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float q = p[i];
    dp[i] = g * (q - target) / fmaxf(q * (1.f - q), 1e-12f);
  }
}

int check() { return cudaGetLastError() == cudaSuccess ? 0 : 4; }


}  // namespace

int nchw_f32_to_nhwc_bf16(const float* x, __nv_bfloat16* y, int N, int C, int H, int W,
                          cudaStream_t s) {
  if (C == 3 || C == 1) {
    const long long npix = static_cast<long long>(N) * H * W;
    if (C == 3)
      nchw_f32_to_nhwc_bf16_small_kernel<3><<<grid_for(npix, kThreads), kThreads, 0, s>>>(x, y, npix, H * W);
    else
      nchw_f32_to_nhwc_bf16_small_kernel<1><<<grid_for(npix, kThreads), kThreads, 0, s>>>(x, y, npix, H * W);
  } else if (C < 32) {
    const long long total = static_cast<long long>(N) * C * H * W;
    nchw_f32_to_nhwc_bf16_kernel<<<grid_for(total, kThreads), kThreads, 0, s>>>(x, y, N, C, H, W);
  } else {
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
    nchw_f32_to_nhwc_bf16_tiled_kernel<<<grid, block, 0, s>>>(x, y, C, H * W);
  }
  return check();
}
int nhwc_bf16_to_nchw_f32(const __nv_bfloat16* x, float* y, int N, int C, int H, int W,
                          cudaStream_t s) {
  if (C == 3) {
    const long long npix = static_cast<long long>(N) * H * W;
    nhwc_bf16_to_nchw_f32_small_kernel<3><<<grid_for(npix, kThreads), kThreads, 0, s>>>(x, y, npix, H * W);
    return check();
  }
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nhwc_bf16_to_nchw_f32_kernel<<<grid, block, 0, s>>>(x, y, C, H * W);
  return check();
}
int tanh_bwd_nchw_to_nhwc(const float* dout, const float* y, __nv_bfloat16* dpre, int N, int C, int H,
                          int W, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * C * H * W;
  tanh_bwd_nchw_to_nhwc_kernel<<<grid_for(total, kThreads), kThreads, 0, s>>>(dout, y, dpre, N, C, H, W);
  return check();
}
int transpose_bf16(const __nv_bfloat16* x, __nv_bfloat16* y, int batch, int R, int Cc,
                   cudaStream_t s) {
  dim3 grid((R + 31) / 32, (Cc + 31) / 32, batch), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, s>>>(x, y, R, Cc);
  return check();
}
int col_stats(const __nv_bfloat16* y, long long M, int C, float* stats, int with_sq, cudaStream_t s) {
  if (C % 8 || C / 8 > kThreads) return 1;
  const int rpb = kThreads / (C / 8);
  col_stats_kernel<<<grid_for(M, rpb, 148 * 4), kThreads, 2 * C * sizeof(float), s>>>(y, M, C, stats,
                                                                                    with_sq);
  return check();
}
int bn_finalize(const float* stats, int stats_rows, float count, const float* gamma, const float* beta,
                float* running_mean, float* running_var, long long* num_batches, float momentum,
                float eps, int training, float* scale, float* shift, float* mean, float* invstd, int C,
                cudaStream_t s) {
  bn_finalize_kernel<<<(C + 7) / 8, kThreads, 0, s>>>(stats, stats_rows, count, gamma, beta, running_mean,
                                                     running_var, num_batches, momentum, eps, training,
                                                     scale, shift, mean, invstd, C);
  return check();
}
int bn_apply(const __nv_bfloat16* y, const float* scale, const float* shift, int act, float slope,
             const float* slope_ptr, const __nv_bfloat16* residual, __nv_bfloat16* out, long long M,
             int C, cudaStream_t s) {
  if (C % 8) return 1;
  const long long nvec = M * C / 8;
  bn_apply_kernel<<<grid_for(nvec, 4 * kThreads, 148 * 8), kThreads, 0, s>>>(y, scale, shift, act, slope, slope_ptr,
                                                                residual, out, nvec, C);
  return check();
}
int bn_bwd_reduce(const __nv_bfloat16* dout, const __nv_bfloat16* y, const float* mean,
                  const float* invstd, const float* scale, const float* shift, int act, float slope,
                  const float* slope_ptr, float* sums, long long M, int C, cudaStream_t s,
                  const PeerTable* peer, int slot, float* sums_global, unsigned int* ticket) {
  if (C % 8 || C / 8 > kThreads) return 1;
  const int rpb = kThreads / (C / 8);
  PeerTable none{};
  none.world = 1;
  if (peer && peer->world > 1 && (!sums_global || !ticket || 2 * C + 1 > kPeerSlotFloats)) return 1;
  bn_bwd_reduce_kernel<<<grid_for(M, rpb * 4, 148 * 2), kThreads, 0, s>>>(
      dout, y, mean, invstd, scale, shift, act, slope, slope_ptr, sums, M, C, peer ? *peer : none, slot,
      sums_global, ticket);
  return check();
}
int bn_bwd_apply(const __nv_bfloat16* dout, const __nv_bfloat16* y, const float* mean,
                 const float* invstd, const float* scale, const float* shift, int act, float slope,
                 const float* slope_ptr, const float* sums, float count, __nv_bfloat16* dy,
                 float* colsum, long long M, int C, cudaStream_t s) {
  if (C % 8 || C / 8 > kThreads) return 1;
  const int rpb = kThreads / (C / 8);
  bn_bwd_apply_kernel<<<grid_for(M, rpb * 4, 148 * 4), kThreads, 0, s>>>(
      dout, y, mean, invstd, scale, shift, act, slope, slope_ptr, sums, 1.f / count, dy, colsum, M, C);
  return check();
}
int act_bwd(const __nv_bfloat16* dout, const __nv_bfloat16* out, int act, float slope,
            const float* slope_ptr, __nv_bfloat16* din, float* dslope, float* colsum, long long M, int C,
            cudaStream_t s) {
  if (C % 8 || C / 8 > kThreads) return 1;
  const int rpb = kThreads / (C / 8);
  act_bwd_kernel<<<grid_for(M, rpb * 4, 148 * 4), kThreads, 0, s>>>(dout, out, act, slope, slope_ptr, din,
                                                                    dslope, colsum, M, C);
  return check();
}
int maxpool2_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, int N, int H, int W, int C,
                 cudaStream_t s) {
  if (C % 8 || H % 2 || W % 2) return 1;
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  maxpool2_fwd_kernel<<<grid_for(total, kThreads), kThreads, 0, s>>>(x, y, N, H, W, C);
  return check();
}
int maxpool2_bwd(const __nv_bfloat16* x, const __nv_bfloat16* dy, __nv_bfloat16* dx, int N, int H,
                 int W, int C, cudaStream_t s) {
  if (C % 8 || H % 2 || W % 2) return 1;
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  maxpool2_bwd_kernel<<<grid_for(total, kThreads), kThreads, 0, s>>>(x, dy, dx, N, H, W, C);
  return check();
}
int pixel_shuffle2(const __nv_bfloat16* src, __nv_bfloat16* dst, int N, int H, int W, int C, int inverse,
                   cudaStream_t s) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0) return 1;
  const long long total = static_cast<long long>(N) * H * W * 4 * C;
  pixel_shuffle2_kernel<<<grid_for(total, kThreads), kThreads, 0, s>>>(src, dst, N, H, W, C, inverse);
  return check();
}
int lr_from_hr(const float* hr, float* lr, const float* dlr, float* dhr, int N, int C, int H, int W, int OH,
               int OW, cudaStream_t s) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) return 1;
  const float sh = OH > 1 ? static_cast<float>(H - 1) / static_cast<float>(OH - 1) : 0.f;
  const float sw = OW > 1 ? static_cast<float>(W - 1) / static_cast<float>(OW - 1) : 0.f;
  if (dlr) cudaMemsetAsync(dhr, 0, sizeof(float) * static_cast<size_t>(N) * C * H * W, s);
  const long long total = static_cast<long long>(N) * C * OH * OW;
  lr_from_hr_kernel<<<grid_for(total, kThreads), kThreads, 0, s>>>(hr, lr, dlr, dhr, N * C, H, W, OH, OW, sh, sw);
  return check();
}
int mse_fwd(const float* a, const float* b, long long n, float coef, float* loss, cudaStream_t s) {
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  mse_fwd_kernel<<<grid_for(n, kThreads, 148 * 4), kThreads, 0, s>>>(a, b, n, coef, loss);
  return check();
}
int mse_bwd(const float* a, const float* b, long long n, float coef, const float* gout, float* ga,
            float* gb, cudaStream_t s) {
  mse_bwd_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(a, b, n, coef, gout, ga, gb);
  return check();
}
int bce_fwd(const float* p, int n, float target, float* loss, float* mean_p, cudaStream_t s) {
  bce_kernel<<<1, kThreads, 0, s>>>(p, n, target, loss, mean_p);
  return check();
}
int bce_bwd(const float* p, int n, float target, const float* gout, float* dp, cudaStream_t s) {
  bce_bwd_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(p, n, target, gout, dp);
  return check();
}

}  // namespace sisr
