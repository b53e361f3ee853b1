// Discriminator head: Linear(fc_in -> fc_mid) + LeakyReLU + Linear(fc_mid -> 1) + Sigmoid,
// forward and backward.  The three skinny GEMMs (batch x 18432 x 1024) are HBM-bound on the 75 MB
// fp32 master weight, which is read in its native [fc_mid, C*H*W] layout (the activations are
// permuted to the reference's (c,h,w) flatten order instead of permuting the 19 M-element weight)
// and rounded to bf16 on the way into shared memory; the contraction is warp-level mma.sync.
#include "linear.h"

#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

namespace sisr {

namespace {

constexpr int kTile = 64;      // block tile is 64 x 64 x 64
constexpr int kLd = kTile + 8; // bf16 row stride of a staged tile (144 B: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                              uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16 consecutive elements of one source row -> 16 bf16 (8 packed words); zero outside [0, valid)
struct Row16 {
  uint32_t w[8];
};
__device__ __forceinline__ Row16 load_row16(const float* p, int valid) {
  Row16 r;
  if (valid >= 16 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(p4 + i);
      r.w[2 * i] = pack_bf16x2(v.x, v.y);
      r.w[2 * i + 1] = pack_bf16x2(v.z, v.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.w[i] = pack_bf16x2(2 * i < valid ? p[2 * i] : 0.f, 2 * i + 1 < valid ? p[2 * i + 1] : 0.f);
  }
  return r;
}
__device__ __forceinline__ Row16 load_row16(const __nv_bfloat16* p, int valid) {
  Row16 r;
  if (valid >= 16 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    r.w[0] = u0.x; r.w[1] = u0.y; r.w[2] = u0.z; r.w[3] = u0.w;
    r.w[4] = u1.x; r.w[5] = u1.y; r.w[6] = u1.z; r.w[7] = u1.w;
  } else {
    const unsigned short* q = reinterpret_cast<const unsigned short*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t lo = 2 * i < valid ? q[2 * i] : 0u, hi = 2 * i + 1 < valid ? q[2 * i + 1] : 0u;
      r.w[i] = lo | (hi << 16);
    }
  }
  return r;
}

// C[M,N] (+)= sum_k A(m,k) * B(n,k) on warp-level mma.sync tiles (bf16 inputs, fp32 accumulate);
// fp32 operands are rounded to bf16 while they are staged into shared memory.
//   AK: A(m,k) = a[m*lda + k]  (k contiguous)   else A(m,k) = a[k*lda + m]
//   BK: B(n,k) = b[n*ldb + k]                   else B(n,k) = b[k*ldb + n]
// The staged tiles keep the source's contiguous dimension, ldmatrix(.trans) builds the fragments.
// gridDim.z > 1: split-K with fp32 atomics into a zeroed C.
template <typename TA, typename TB, bool AK, bool BK>
__global__ void __launch_bounds__(256)
gemm_mma_kernel(const TA* __restrict__ a, long long lda, const TB* __restrict__ b, long long ldb,
                float* __restrict__ c, long long ldc, int M, int N, int K, int k_per_split) {
  __shared__ __align__(16) __nv_bfloat16 As[2][kTile * kLd];
  __shared__ __align__(16) __nv_bfloat16 Bs[2][kTile * kLd];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int lrow = tid >> 2, lseg = (tid & 3) * 16;
  auto fetch_a = [&](int k0) {
    if (AK) {
      const int m = m0 + lrow, k = k0 + lseg;
      return load_row16(a + static_cast<long long>(m) * lda + k, m < M ? kend - k : 0);
    }
    const int k = k0 + lrow, m = m0 + lseg;
    return load_row16(a + static_cast<long long>(k) * lda + m, k < kend ? M - m : 0);
  };
  auto fetch_b = [&](int k0) {
    if (BK) {
      const int n = n0 + lrow, k = k0 + lseg;
      return load_row16(b + static_cast<long long>(n) * ldb + k, n < N ? kend - k : 0);
    }
    const int k = k0 + lrow, n = n0 + lseg;
    return load_row16(b + static_cast<long long>(k) * ldb + n, k < kend ? N - n : 0);
  };
  auto stash = [&](__nv_bfloat16* dst, const Row16& r) {
    uint4* d = reinterpret_cast<uint4*>(dst + lrow * kLd + lseg);
    d[0] = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
    d[1] = make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]);
  };
  const int wm = warp & 3, wn = warp >> 2;
  float acc[4][4] = {};
  Row16 ra = fetch_a(kbeg), rb = fetch_b(kbeg);
  stash(As[0], ra);
  stash(Bs[0], rb);
  __syncthreads();
  int cur = 0;
  for (int k0 = kbeg; k0 < kend; k0 += kTile) {
    const bool more = k0 + kTile < kend;
    if (more) {
      ra = fetch_a(k0 + kTile);
      rb = fetch_b(k0 + kTile);
    }
    const __nv_bfloat16* at = As[cur];
    const __nv_bfloat16* bt = Bs[cur];
#pragma unroll
    for (int ks = 0; ks < kTile / 16; ++ks) {
      uint32_t a0, a1, a2, a3;
      if (AK)
        ldsm_x4(smem_u32(at + (wm * 16 + (lane & 15)) * kLd + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
      else
        ldsm_x4_trans(smem_u32(at + (ks * 16 + (lane & 7) + ((lane >> 4) << 3)) * kLd + wm * 16 +
                               ((lane >> 3) & 1) * 8),
                      a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        if (BK)
          ldsm_x4(smem_u32(bt + (wn * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3)) * kLd + ks * 16 +
                           ((lane >> 3) & 1) * 8),
                  b0, b1, b2, b3);
        else
          ldsm_x4_trans(smem_u32(bt + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kLd + wn * 32 +
                                 np * 16 + ((lane >> 4) << 3)),
                        b0, b1, b2, b3);
        mma_bf16(acc[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    if (more) {
      stash(As[cur ^ 1], ra);
      stash(Bs[cur ^ 1], rb);
    }
    __syncthreads();
    cur ^= 1;
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gm = m0 + wm * 16 + g + (e >> 1) * 8;
      const int gn = n0 + wn * 32 + nt * 8 + 2 * t + (e & 1);
      if (gm >= M || gn >= N) continue;
      float* dst = c + gm * ldc + gn;
      if (gridDim.z > 1) atomicAdd(dst, acc[nt][e]); else *dst = acc[nt][e];
    }
}

template <typename TA, typename TB, bool AK, bool BK>
int launch_gemm(const TA* a, long long lda, const TB* b, long long ldb, float* c, long long ldc, int M,
                int N, int K, cudaStream_t s) {
  const int tiles = ((M + kTile - 1) / kTile) * ((N + kTile - 1) / kTile);
  int splits = 1;
  if (tiles < 148 && K >= 1024) {
    splits = (296 + tiles - 1) / tiles;
    const int max_splits = K / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = (K + splits - 1) / splits;
  kps = (kps + kTile - 1) / kTile * kTile;
  splits = (K + kps - 1) / kps;
  if (splits > 1) cudaMemsetAsync(c, 0, sizeof(float) * static_cast<size_t>(M) * ldc, s);
  dim3 grid((N + kTile - 1) / kTile, (M + kTile - 1) / kTile, splits);
  gemm_mma_kernel<TA, TB, AK, BK><<<grid, 256, 0, s>>>(a, lda, b, ldb, c, ldc, M, N, K, kps);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

// ------------------------------------------------------------------ weight-streaming GEMMs of the head
// The three GEMMs above move 75 MB each through a one-tile register prefetch and stop at 1.5-2.6 TB/s.
// The kernels below keep the same mma.sync contraction but stream the operands through a 3-stage
// cp.async pipeline as RAW bytes (fp32 stays fp32 in shared memory and is rounded to bf16 when a
// fragment is built), two blocks per SM, so ~64 KB of weight bytes are in flight per SM.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kWsStages = 3;
// C[M, N] (+)= A[M, K] * W   with M <= 64 per block row, W fp32 streamed once:
//   BK = true : W(n, k) = w[n * ldw + k]   (head forward:  h = x . W0^T,  A = x bf16)
//   BK = false: W(k, n) = w[k * ldw + n]   (head dgrad:    dx = dh . W0,  A = dh fp32)
// K % 64 == 0, N % 64 == 0.  gridDim = (N/64, ceil(M/64), splits); splits > 1: fp32 atomics into zeroed C.
template <typename TA, bool BK>
__global__ void __launch_bounds__(256, 2)
wstream_gemm_kernel(const TA* __restrict__ a, long long lda, const float* __restrict__ w, long long ldw,
                    float* __restrict__ c, long long ldc, int M, int K, int k_per_split) {
  constexpr bool A16 = sizeof(TA) == 2;
  constexpr int PW = BK ? 72 : 68;                    // fp32 pitch of the W tile (conflict-free fragment reads)
  constexpr int PA = 72;                              // pitch of the A tile in elements
  constexpr int kWBytes = 64 * PW * 4;
  constexpr int kABytes = 64 * PA * static_cast<int>(sizeof(TA));
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int nk = (kend - kbeg) / 64;
  auto w_tile = [&](int st) { return reinterpret_cast<float*>(smem_raw + st * (kWBytes + kABytes)); };
  auto a_tile = [&](int st) { return reinterpret_cast<TA*>(smem_raw + st * (kWBytes + kABytes) + kWBytes); };

  auto load_stage = [&](int st, int kt) {
    const int k0 = kbeg + kt * 64;
    float* ws = w_tile(st);
#pragma unroll
    for (int i = 0; i < 4; ++i) {                     // 64 rows x 16 chunks of 16 B
      const int idx = tid + i * 256, row = idx >> 4, ch = idx & 15;
      const float* src = BK ? w + static_cast<long long>(n0 + row) * ldw + k0 + ch * 4
                            : w + static_cast<long long>(k0 + row) * ldw + n0 + ch * 4;
      cp_async16(smem_u32(ws + row * PW + ch * 4), src);
    }
    TA* as = a_tile(st);
    constexpr int kChunks = A16 ? 8 : 16;             // 16-byte chunks per 64-element row
#pragma unroll
    for (int i = 0; i < kChunks / 4; ++i) {
      const int idx = tid + i * 256, row = idx / kChunks, ch = idx - row * kChunks;
      TA* dst = as + row * PA + ch * (16 / static_cast<int>(sizeof(TA)));
      if (m0 + row < M)
        cp_async16(smem_u32(dst), a + static_cast<long long>(m0 + row) * lda + k0 + ch * (16 / static_cast<int>(sizeof(TA))));
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  };

  float acc[4][4] = {};
#pragma unroll
  for (int s_ = 0; s_ < kWsStages - 1; ++s_) {
    if (s_ < nk) load_stage(s_, s_);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kWsStages - 2>();
    __syncthreads();
    if (kt + kWsStages - 1 < nk) load_stage((kt + kWsStages - 1) % kWsStages, kt + kWsStages - 1);
    cp_async_commit();
    const float* ws = w_tile(kt % kWsStages);
    const TA* as = a_tile(kt % kWsStages);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a0, a1, a2, a3;
      if constexpr (A16) {
        ldsm_x4(smem_u32(as + (wm * 16 + (lane & 15)) * PA + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
      } else {
        const float* ar = reinterpret_cast<const float*>(as) + (wm * 16 + g) * PA + ks * 16 + 2 * t;
        const float2 f0 = *reinterpret_cast<const float2*>(ar);
        const float2 f1 = *reinterpret_cast<const float2*>(ar + 8 * PA);
        const float2 f2 = *reinterpret_cast<const float2*>(ar + 8);
        const float2 f3 = *reinterpret_cast<const float2*>(ar + 8 * PA + 8);
        a0 = pack_bf16x2(f0.x, f0.y); a1 = pack_bf16x2(f1.x, f1.y);
        a2 = pack_bf16x2(f2.x, f2.y); a3 = pack_bf16x2(f3.x, f3.y);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn * 32 + nt * 8 + g;
        uint32_t b0, b1;
        if constexpr (BK) {
          const float* br = ws + n * PW + ks * 16 + 2 * t;
          const float2 f0 = *reinterpret_cast<const float2*>(br);
          const float2 f1 = *reinterpret_cast<const float2*>(br + 8);
          b0 = pack_bf16x2(f0.x, f0.y);
          b1 = pack_bf16x2(f1.x, f1.y);
        } else {
          const float* br = ws + (ks * 16 + 2 * t) * PW + n;
          b0 = pack_bf16x2(br[0], br[PW]);
          b1 = pack_bf16x2(br[8 * PW], br[9 * PW]);
        }
        mma_bf16(acc[nt], a0, a1, a2, a3, b0, b1);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int gm = m0 + wm * 16 + g + hh * 8;
      const int gn = n0 + wn * 32 + nt * 8 + 2 * t;
      if (gm >= M) continue;
      float* dst = c + gm * ldc + gn;
      if (gridDim.z > 1) {
        atomicAdd(dst, acc[nt][hh * 2]);
        atomicAdd(dst + 1, acc[nt][hh * 2 + 1]);
      } else {
        *reinterpret_cast<float2*>(dst) = make_float2(acc[nt][hh * 2], acc[nt][hh * 2 + 1]);
      }
    }
}

template <typename TA, bool BK>
int launch_wstream(const TA* a, long long lda, const float* w, long long ldw, float* c, long long ldc, int M,
                   int N, int K, cudaStream_t s) {
  constexpr int PW = BK ? 72 : 68;
  constexpr int smem = kWsStages * (64 * PW * 4 + 64 * 72 * static_cast<int>(sizeof(TA)));
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wstream_gemm_kernel<TA, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return 4;
    configured = true;
  }
  const int tiles = (N / 64) * ((M + 63) / 64);
  int splits = 1;
  if (tiles < 148) {
    splits = (148 * 2 + tiles - 1) / tiles;
    if (splits > K / 256) splits = K / 256;
    if (splits < 1) splits = 1;
  }
  int kps = ((K / 64 + splits - 1) / splits) * 64;
  splits = (K + kps - 1) / kps;
  if (splits > 1) cudaMemsetAsync(c, 0, sizeof(float) * static_cast<size_t>(M) * ldc, s);
  dim3 grid(N / 64, (M + 63) / 64, splits);
  wstream_gemm_kernel<TA, BK><<<grid, 256, smem, s>>>(a, lda, w, ldw, c, ldc, M, K, kps);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

// dW[m, n] = sum_{b < B} dh[b, m] * x[b, n]   (B <= 64: one k-tile; fp32 output written once, 75 MB)
// One block = 64 rows m (A fragments of dh^T live in registers) x a strided set of 64-column tiles of x,
// streamed through a 3-stage cp.async pipeline.  gridDim = (column groups, mid / 64).
__global__ void __launch_bounds__(256, 2)
dhead_wgrad_kernel(const float* __restrict__ dh, int mid, const __nv_bfloat16* __restrict__ x, long long ldx,
                   float* __restrict__ dw, long long ldw, int B, int n_tiles) {
  constexpr int PD = 68;                              // fp32 pitch of the dh tile [k = sample][m]
  __shared__ __align__(16) float dhs[64 * PD];
  __shared__ __align__(16) __nv_bfloat16 xs[kWsStages][64 * kLd];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int m0 = blockIdx.y * 64;
  for (int i = tid; i < 64 * 16; i += 256) {          // dh tile: 64 samples x 64 features
    const int row = i >> 4, ch = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < B) v = *reinterpret_cast<const float4*>(dh + static_cast<long long>(row) * mid + m0 + ch * 4);
    *reinterpret_cast<float4*>(dhs + row * PD + ch * 4) = v;
  }
  auto load_x = [&](int st, int tile) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {                     // 64 samples x 8 chunks of 16 B
      const int idx = tid + i * 256, row = idx >> 3, ch = idx & 7;
      __nv_bfloat16* dst = xs[st] + row * kLd + ch * 8;
      if (row < B)
        cp_async16(smem_u32(dst), x + static_cast<long long>(row) * ldx + static_cast<long long>(tile) * 64 + ch * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  };
  const int first = blockIdx.x, step = gridDim.x;
  const int mine = first < n_tiles ? (n_tiles - first + step - 1) / step : 0;
#pragma unroll
  for (int s_ = 0; s_ < kWsStages - 1; ++s_) {
    if (s_ < mine) load_x(s_, first + s_ * step);
    cp_async_commit();
  }
  __syncthreads();
  uint32_t af[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float* ar = dhs + (ks * 16 + 2 * t) * PD + wm * 16 + g;     // A(m, k) = dh[k][m]
    af[ks][0] = pack_bf16x2(ar[0], ar[PD]);
    af[ks][1] = pack_bf16x2(ar[8], ar[PD + 8]);
    af[ks][2] = pack_bf16x2(ar[8 * PD], ar[9 * PD]);
    af[ks][3] = pack_bf16x2(ar[8 * PD + 8], ar[9 * PD + 8]);
  }
  for (int it = 0; it < mine; ++it) {
    cp_async_wait<kWsStages - 2>();
    __syncthreads();
    if (it + kWsStages - 1 < mine) load_x((it + kWsStages - 1) % kWsStages, first + (it + kWsStages - 1) * step);
    cp_async_commit();
    const __nv_bfloat16* bt = xs[it % kWsStages];
    float acc[4][4] = {};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(smem_u32(bt + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kLd + wn * 32 + np * 16 +
                               ((lane >> 4) << 3)),
                      b0, b1, b2, b3);
        mma_bf16(acc[2 * np], af[ks][0], af[ks][1], af[ks][2], af[ks][3], b0, b1);
        mma_bf16(acc[2 * np + 1], af[ks][0], af[ks][1], af[ks][2], af[ks][3], b2, b3);
      }
    const long long nb = static_cast<long long>(first + it * step) * 64;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float* dst = dw + static_cast<long long>(m0 + wm * 16 + g + hh * 8) * ldw + nb + wn * 32 + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(dst) = make_float2(acc[nt][hh * 2], acc[nt][hh * 2 + 1]);
      }
  }
  cp_async_wait<0>();
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

// (the generic gemm_mma_kernel remains the fallback for head sizes the streaming kernels do not take)
constexpr bool head_v1() { return false; }

}  // namespace

int dhead_forward(const __nv_bfloat16* x_flat, const float* w0, const float* b0, const float* w2,
                  const float* b2, float slope, float* h, float* p, int B, int fc_in, int fc_mid,
                  cudaStream_t s) {
  const bool stream_ok = fc_in % 64 == 0 && fc_mid % 64 == 0 && !head_v1() &&
                         (reinterpret_cast<uintptr_t>(w0) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_flat) & 15) == 0;
  if (stream_ok) {
    if (int rc = launch_wstream<__nv_bfloat16, true>(x_flat, fc_in, w0, fc_in, h, fc_mid, B, fc_mid, fc_in, s))
      return rc;
  } else if (int rc = launch_gemm<__nv_bfloat16, float, true, true>(x_flat, fc_in, w0, fc_in, h, fc_mid, B,
                                                                    fc_mid, fc_in, s)) {
    return rc;
  }
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
  const bool stream_ok = fc_in % 64 == 0 && fc_mid % 64 == 0 && !head_v1() &&
                         (reinterpret_cast<uintptr_t>(w0) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_flat) & 15) == 0;
  if (need_wgrad) {
    // dW0[m=fc_mid, n=fc_in] = sum_b dh[b,m] * x[b,n]
    if (stream_ok && B <= 64 && (reinterpret_cast<uintptr_t>(dw0) & 7) == 0) {
      const int n_tiles = fc_in / 64;
      dim3 grid(n_tiles < 37 ? n_tiles : 37, fc_mid / 64);
      dhead_wgrad_kernel<<<grid, 256, 0, s>>>(dh, fc_mid, x_flat, fc_in, dw0, fc_in, B, n_tiles);
    } else if (int rc = launch_gemm<float, __nv_bfloat16, false, false>(dh, fc_mid, x_flat, fc_in, dw0, fc_in,
                                                                        fc_mid, fc_in, B, s)) {
      return rc;
    }
  }
  if (dx_flat) {
    // dx[b, n=fc_in] = sum_k dh[b,k] * W0[k,n]
    if (stream_ok) {
      if (int rc = launch_wstream<float, false>(dh, fc_mid, w0, fc_in, dx_flat, fc_in, B, fc_in, fc_mid, s))
        return rc;
    } else if (int rc = launch_gemm<float, float, true, false>(dh, fc_mid, w0, fc_in, dx_flat, fc_in, B, fc_in,
                                                                fc_mid, s)) {
      return rc;
    }
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

}  // namespace sisr
