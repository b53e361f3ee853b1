// Kernels for the "thin" edge layers of the SRGAN step, where one side of the convolution has
// 3 channels (RGB):
//   thin-in   : 3 -> CW   (G first conv 9x9, D / VGG first conv 3x3, dgrad of the G output conv)
//   thin-out  : CW -> 3   (G output conv 3x3 + Tanh, dgrad of the D / VGG first conv)
//   thin-wgrad: correlation of a 3-channel tensor with a wide tensor over the k x k taps
// Stride 1, same-size.  These layers are < 2 % of the step's FLOPs but stream the largest
// activations of the step (B x 96 x 96 x 64), so they are HBM-bound: each kernel stages its tile in
// shared memory with 16-byte accesses and contracts it with warp-level mma.sync (m16n8k16, bf16
// inputs, fp32 accumulate) -- a tcgen05 tile (N >= 16, K = 64 per swizzle atom) would be > 80 %
// padding on a K = 27 or N = 3 problem.
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
// D(16x8, fp32) += A(16x16, row) * B(16x8, col), bf16 inputs
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

constexpr int kThreads = 256;
constexpr int kTileQ = 128;    // pixels per tile (thin-in, wgrad)
constexpr int kLdC = 72;       // bf16 row stride of a staged [pixels][64 channels] tile (144 B)

// Stage the pixels [w0, w0 + count) (flattened N*H*W order, 3 bf16 each, contiguous in memory) of a
// 3-channel tensor into shared memory with 16-byte loads; pixels outside [0, npix) are left unset
// (the im2col below only reads pixels of valid taps).  Returns nothing; win[(q - w0) * 3 + c].
__device__ __forceinline__ void stage_window3(const __nv_bfloat16* __restrict__ src, long long npix,
                                              long long w0, int count, __nv_bfloat16* win) {
  // byte range of the window, widened to 16-byte boundaries of the global buffer
  const long long lo = (w0 < 0 ? 0 : w0) * 6;
  long long hi = (w0 + count) * 6;
  if (hi > npix * 6) hi = npix * 6;
  if (hi <= lo) return;
  const uintptr_t gbase = reinterpret_cast<uintptr_t>(src);
  const long long a0 = static_cast<long long>(((gbase + lo) & ~static_cast<uintptr_t>(15)) - gbase);
  // destination byte offset of global byte a0 inside `win` is (a0 - w0*6); may be negative by < 16:
  // the window buffer has 16 bytes of slack in front (see callers)
  const long long tot = static_cast<long long>(npix) * 6;
  for (long long a = a0 + threadIdx.x * 16LL; a < hi; a += blockDim.x * 16LL) {
    const char* g = reinterpret_cast<const char*>(src) + a;
    char* d = reinterpret_cast<char*>(win) + (a - w0 * 6);
    if (a >= 0 && a + 16 <= tot) {
      const uint4 v = *reinterpret_cast<const uint4*>(g);
      // shared destination is only 2-byte aligned in general: store as 8 halves
      const unsigned short* h = reinterpret_cast<const unsigned short*>(&v);
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<unsigned short*>(d)[i] = h[i];
    } else {
      for (int i = 0; i < 16 && a + i < tot; i += 2)
        if (a + i >= 0)
          *reinterpret_cast<unsigned short*>(d + i) = *reinterpret_cast<const unsigned short*>(g + i);
    }
  }
}

// ------------------------------------------------------------------ thin-in: 3 -> 64, k x k
// GEMM per 128-pixel tile: Y[128, 64] = Xim[128, KP] * W[64, KP]^T, KP = k*k*3 zero-padded.
// w: [CW][k*k][3] bf16; flip: the tap at position t uses w[.][k*k-1-t][.] (transposed conv).
template <int KP>
__global__ void __launch_bounds__(kThreads)
thin_in_mma_kernel(ThinConv c, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                   const float* __restrict__ bias, int act, float slope,
                   const float* __restrict__ slope_ptr, int flip, __nv_bfloat16* __restrict__ y) {
  constexpr int LDK = KP + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [64][LDK]
  __nv_bfloat16* Xs = Ws + 64 * LDK;                                // [128][LDK]
  __nv_bfloat16* Cs = Xs + kTileQ * LDK;                            // [128][kLdC]
  __nv_bfloat16* Win = Cs + kTileQ * kLdC + 8;                      // input window (+16 B slack in front)
  const int T = c.k * c.k, KT = T * 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cwb = blockIdx.y * 64;
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  const long long q0 = static_cast<long long>(blockIdx.x) * kTileQ;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);

  for (int i = tid; i < 64 * LDK; i += kThreads) {
    const int cw = i / LDK, j = i - cw * LDK;
    __nv_bfloat16 v = zero;
    if (j < KT) {
      const int tap = j / 3, cs = j - tap * 3;
      const int st = flip ? T - 1 - tap : tap;
      v = w[(static_cast<size_t>(cwb + cw) * T + st) * 3 + cs];
    }
    Ws[i] = v;
  }
  const int reach = c.pad * (c.W + 1);                 // farthest tap in flattened pixels
  const long long win0 = q0 - reach;
  stage_window3(x, npix, win0, kTileQ + 2 * reach, Win);
  __syncthreads();
  {  // im2col of the tile from the staged window: two threads per pixel, alternating taps
    const int ql = tid >> 1, half = tid & 1;
    const long long q = q0 + ql;
    const bool valid = q < npix;
    const long long qq = valid ? q : 0;
    const int ow = static_cast<int>(qq % c.W);
    const int oh = static_cast<int>((qq / c.W) % c.H);
    __nv_bfloat16* row = Xs + ql * LDK;
    for (int tap = half; tap < T; tap += 2) {
      const int kh = tap / c.k, kw = tap - kh * c.k;
      const int ih = oh - c.pad + kh, iw = ow - c.pad + kw;
      __nv_bfloat16 v0 = zero, v1 = zero, v2 = zero;
      if (valid && ih >= 0 && ih < c.H && iw >= 0 && iw < c.W) {
        const __nv_bfloat16* p = Win + (ql + reach + (kh - c.pad) * c.W + (kw - c.pad)) * 3;
        v0 = p[0]; v1 = p[1]; v2 = p[2];
      }
      row[tap * 3] = v0; row[tap * 3 + 1] = v1; row[tap * 3 + 2] = v2;
    }
    for (int j = KT + half; j < KP; j += 2) row[j] = zero;
  }
  __syncthreads();

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int r0 = warp * 16;
  const uint32_t a_base = smem_u32(Xs + (r0 + (lane & 15)) * LDK + (lane >> 4) * 8);
  const uint32_t b_base = smem_u32(Ws + ((lane & 7) + ((lane >> 4) << 3)) * LDK + ((lane >> 3) & 1) * 8);
#pragma unroll 4
  for (int ks = 0; ks < KP / 16; ++ks) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(a_base + ks * 32, a0, a1, a2, a3);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b_base + (np * 16 * LDK) * 2 + ks * 32, b0, b1, b2, b3);
      mma_bf16(acc[2 * np], a0, a1, a2, a3, b0, b1);
      mma_bf16(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  if (act == ACT_PRELU) slope = *slope_ptr;
  if (act == ACT_RELU) slope = 0.f;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * t;
    const float b0 = bias ? bias[cwb + col] : 0.f, b1 = bias ? bias[cwb + col + 1] : 0.f;
    *reinterpret_cast<uint32_t*>(Cs + (r0 + g) * kLdC + col) =
        pack_bf16x2(act_apply(acc[nt][0] + b0, act, slope), act_apply(acc[nt][1] + b1, act, slope));
    *reinterpret_cast<uint32_t*>(Cs + (r0 + g + 8) * kLdC + col) =
        pack_bf16x2(act_apply(acc[nt][2] + b0, act, slope), act_apply(acc[nt][3] + b1, act, slope));
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = tid + it * kThreads;
    const int row = idx >> 3, chunk = idx & 7;
    if (q0 + row < npix)
      *reinterpret_cast<uint4*>(y + (q0 + row) * c.CW + cwb + chunk * 8) =
          *reinterpret_cast<const uint4*>(Cs + row * kLdC + chunk * 8);
  }
}

// ------------------------------------------------------------------ thin-out: 64 -> 3, 3 x 3
// One block = R output rows of one image; the (R+2) x (W+2) x 64 input halo tile is staged once and
// every 16-pixel group is a [16 x 576] x [576 x 8] product (3 output channels padded to 8).
// w: [3][9][64] bf16; y_bf16: [N,H,W,3] and/or y_nchw: [N,3,H,W] fp32
__global__ void __launch_bounds__(kThreads)
thin_out_mma_kernel(ThinConv c, int R, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias, int act,
                    float slope, int flip, __nv_bfloat16* __restrict__ y_bf16,
                    float* __restrict__ y_nchw) {
  constexpr int LDW = 9 * 64 + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [8][LDW]
  __nv_bfloat16* Xs = Ws + 8 * LDW;                                 // [(R+2)*(W+2)][kLdC]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, row0 = blockIdx.x * R;
  const int WP = c.W + 2;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int i = tid; i < 8 * LDW; i += kThreads) {
    const int cs = i / LDW, j = i - cs * LDW;
    __nv_bfloat16 v = zero;
    if (cs < 3 && j < 576) {
      const int tap = j >> 6, cw = j & 63;
      const int st = flip ? 8 - tap : tap;
      v = w[(static_cast<size_t>(cs) * 9 + st) * 64 + cw];
    }
    Ws[i] = v;
  }
  const int npos = (R + 2) * WP;
  const __nv_bfloat16* xn = x + static_cast<size_t>(n) * c.H * c.W * 64;
  for (int i = tid; i < npos * 8; i += kThreads) {
    const int pix = i >> 3, ch = i & 7;
    const int pr = pix / WP, pc = pix - pr * WP;
    const int ih = row0 - 1 + pr, iw = pc - 1;
    __nv_bfloat16* dst = Xs + pix * kLdC + ch * 8;
    if (ih >= 0 && ih < c.H && iw >= 0 && iw < c.W)
      cp_async16(smem_u32(dst), xn + (static_cast<size_t>(ih) * c.W + iw) * 64 + ch * 8);
    else
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  cp_async_wait_all();
  __syncthreads();

  const int npx = R * c.W;
  const int ntiles = (npx + 15) >> 4;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t b_base = smem_u32(Ws + (lane & 7) * LDW + (lane >> 3) * 8);
  for (int mt = warp; mt < ntiles; mt += kThreads / 32) {
    int pl = mt * 16 + (lane & 15);
    if (pl >= npx) pl = npx - 1;
    const int pr = pl / c.W, pc = pl - pr * c.W;
    const uint32_t a_base = smem_u32(Xs + (pr * WP + pc) * kLdC + (lane >> 4) * 8);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int kh = tap / 3, kw = tap - kh * 3;
      const uint32_t a_tap = a_base + ((kh * WP + kw) * kLdC) * 2;
#pragma unroll
      for (int kc = 0; kc < 4; kc += 2) {
        uint32_t b0, b1, b2, b3, a0, a1, a2, a3;
        ldsm_x4(b_base + (tap * 64 + kc * 16) * 2, b0, b1, b2, b3);
        ldsm_x4(a_tap + kc * 32, a0, a1, a2, a3);
        mma_bf16(acc, a0, a1, a2, a3, b0, b1);
        ldsm_x4(a_tap + kc * 32 + 32, a0, a1, a2, a3);
        mma_bf16(acc, a0, a1, a2, a3, b2, b3);
      }
    }
    if (t < 2) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int p = mt * 16 + g + hh * 8;
        const int oh = row0 + p / c.W, ow = p % c.W;
        if (p >= npx || oh >= c.H) continue;
        const size_t q = (static_cast<size_t>(n) * c.H + oh) * c.W + ow;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cs = 2 * t + e;
          if (cs >= 3) continue;
          const float v = act_apply(acc[hh * 2 + e] + (bias ? bias[cs] : 0.f), act, slope);
          if (y_bf16) y_bf16[q * 3 + cs] = __float2bfloat16_rn(v);
          if (y_nchw) y_nchw[((static_cast<size_t>(n) * 3 + cs) * c.H + oh) * c.W + ow] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ thin wgrad
// G[cw, j] = sum_q wide[q, cw] * small[q + sgn*(tap - pad), cs],  j = tap*3 + cs
//   type A (sgn=+1): small = input x, wide = dy:  out[(cw*T + tap)*3 + cs]
//   type B (sgn=-1): small = dy, wide = input x:  out[(cs*T + tap)*CW + cw]
// GEMM with K = pixels: A = wide^T (ldmatrix.trans from the staged [pixels][64] tile), B = the
// transposed im2col XT[j][pixel] built in shared memory; column j = T*3 of XT is all ones, which
// yields sum_q wide[q, cw] (the bias gradient of a thin-in conv) for free.
// NS = number of 32-column slices of j: 1 (3x3: the 8 warps split the pixels of a tile) or
// 8 (9x9: the 8 warps split j).  fp32 atomics into `out` (zeroed by the caller).
template <int NS>
__global__ void __launch_bounds__(kThreads)
thin_wgrad_mma_kernel(ThinConv c, const __nv_bfloat16* __restrict__ small,
                      const __nv_bfloat16* __restrict__ wide, int sgn, float* __restrict__ out,
                      float* __restrict__ wide_colsum, int ntiles) {
  constexpr int NP = NS * 32;
  constexpr int LDQ = kTileQ + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Wd = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [128][kLdC]
  __nv_bfloat16* XT = Wd + kTileQ * kLdC;                           // [NP][LDQ]
  const int T = c.k * c.k, KT = T * 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cwb = blockIdx.y * 64;
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  for (int j = KT + 1 + (tid >> 7); j < NP; j += 2) XT[j * LDQ + (tid & 127)] = zero;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long q0 = static_cast<long long>(tile) * kTileQ;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * kThreads;
      const int row = idx >> 3, chunk = idx & 7;
      __nv_bfloat16* dst = Wd + row * kLdC + chunk * 8;
      if (q0 + row < npix)
        cp_async16(smem_u32(dst), wide + (q0 + row) * c.CW + cwb + chunk * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    {
      const int ql = tid & 127, half = tid >> 7;
      const long long q = q0 + ql;
      const bool valid = q < npix;
      const long long qq = valid ? q : 0;
      const int qw = static_cast<int>(qq % c.W);
      const int qh = static_cast<int>((qq / c.W) % c.H);
      for (int tap = half; tap < T; tap += 2) {
        const int dh = sgn * (tap / c.k - c.pad), dw = sgn * (tap % c.k - c.pad);
        const int ih = qh + dh, iw = qw + dw;
        __nv_bfloat16 v0 = zero, v1 = zero, v2 = zero;
        if (valid && ih >= 0 && ih < c.H && iw >= 0 && iw < c.W) {
          const __nv_bfloat16* p = small + (qq + dh * c.W + dw) * 3;
          v0 = p[0]; v1 = p[1]; v2 = p[2];
        }
        XT[(tap * 3) * LDQ + ql] = v0;
        XT[(tap * 3 + 1) * LDQ + ql] = v1;
        XT[(tap * 3 + 2) * LDQ + ql] = v2;
      }
      if (half == 0) XT[KT * LDQ + ql] = valid ? one : zero;
    }
    cp_async_wait_all();
    __syncthreads();

    const int ns = NS == 1 ? 0 : warp;
    const int ks_begin = NS == 1 ? warp : 0;
    const int ks_end = NS == 1 ? warp + 1 : kTileQ / 16;
    for (int ks = ks_begin; ks < ks_end; ++ks) {
      uint32_t a[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        ldsm_x4_trans(smem_u32(Wd + (ks * 16 + (lane & 7) + ((lane >> 4) << 3)) * kLdC + mt * 16 +
                               ((lane >> 3) & 1) * 8),
                      a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(XT + (ns * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDQ + ks * 16 +
                         ((lane >> 3) & 1) * 8),
                b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          mma_bf16(acc[mt][2 * np], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
          mma_bf16(acc[mt][2 * np + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
        }
      }
    }
  }

  const int g = lane >> 2, t = lane & 3;
  auto emit = [&](int cw, int j, float v) {
    if (j < KT) {
      const int tap = j / 3, cs = j - tap * 3;
      float* dst = (sgn > 0) ? out + static_cast<size_t>(cwb + cw) * KT + j
                             : out + (static_cast<size_t>(cs) * T + tap) * c.CW + cwb + cw;
      atomicAdd(dst, v);
    } else if (j == KT && wide_colsum) {
      atomicAdd(&wide_colsum[cwb + cw], v);
    }
  };
  if (NS == 1) {
    // combine the 8 warps' partial sums (disjoint pixel ranges) in shared memory
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem_raw);   // [64][33] over the Wd region
    for (int wv = 0; wv < kThreads / 32; ++wv) {
      if (warp == wv) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int m = mt * 16 + g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
              float* r = red + m * 33 + j;
              *r = (wv == 0) ? acc[mt][nt][e] : *r + acc[mt][nt][e];
            }
      }
      __syncthreads();
    }
    for (int i = tid; i < 64 * (KT + 1); i += kThreads) {
      const int cw = i / (KT + 1), j = i - cw * (KT + 1);
      emit(cw, j, red[cw * 33 + j]);
    }
  } else {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          emit(mt * 16 + g + (e >> 1) * 8, warp * 32 + nt * 8 + 2 * t + (e & 1), acc[mt][nt][e]);
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

template <typename K>
int set_smem(K kernel, int bytes, int* configured) {
  if (bytes > *configured) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
      return check("cudaFuncSetAttribute");
    *configured = bytes;
  }
  return 0;
}

constexpr int kSmemLimit = 220 * 1024;
// staged input window of thin_in / thin_wgrad: 128 pixels + the reach of the farthest tap on both sides
int thin_window_bytes(const ThinConv& c) { return (kTileQ + 2 * c.pad * (c.W + 1)) * 6 + 48; }

int thin_out_rows(const ThinConv& c) {
  for (int r = 4; r >= 1; r >>= 1) {
    const int rr = r < c.H ? r : c.H;
    if ((8 * (9 * 64 + 8) + (rr + 2) * (c.W + 2) * kLdC) * 2 <= 100 * 1024 || r == 1) return rr;
  }
  return 1;
}

}  // namespace

const char* thin_last_error() { return g_err; }

bool thin_in_supported(const ThinConv& c) {
  return c.CS == 3 && c.CW % 64 == 0 && (c.k == 3 || c.k == 9) &&
         ((64 + kTileQ) * ((c.k == 3 ? 32 : 256) + 8) + kTileQ * kLdC) * 2 + thin_window_bytes(c) <= kSmemLimit;
}
bool thin_out_supported(const ThinConv& c) {
  return c.CS == 3 && c.CW == 64 && c.k == 3 &&
         (8 * (9 * 64 + 8) + 3 * (c.W + 2) * kLdC) * 2 <= kSmemLimit;
}
bool thin_wgrad_supported(const ThinConv& c) { return c.CS == 3 && c.CW % 64 == 0 && (c.k == 3 || c.k == 9); }

int thin_in_conv(const ThinConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias,
                 int act, float slope, const float* slope_ptr, int flip, __nv_bfloat16* y,
                 cudaStream_t s) {
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  dim3 grid(static_cast<unsigned>((npix + kTileQ - 1) / kTileQ), c.CW / 64);
  if (c.k == 3) {
    constexpr int KP = 32;
    const int smem = ((64 + kTileQ) * (KP + 8) + kTileQ * kLdC) * 2 + thin_window_bytes(c);
    static int configured = 0;
    if (int rc = set_smem(thin_in_mma_kernel<KP>, smem, &configured)) return rc;
    thin_in_mma_kernel<KP><<<grid, kThreads, smem, s>>>(c, x, w, bias, act, slope, slope_ptr, flip, y);
  } else {
    constexpr int KP = 256;
    const int smem = ((64 + kTileQ) * (KP + 8) + kTileQ * kLdC) * 2 + thin_window_bytes(c);
    static int configured = 0;
    if (int rc = set_smem(thin_in_mma_kernel<KP>, smem, &configured)) return rc;
    thin_in_mma_kernel<KP><<<grid, kThreads, smem, s>>>(c, x, w, bias, act, slope, slope_ptr, flip, y);
  }
  return check("thin_in_conv");
}

int thin_out_conv(const ThinConv& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias,
                  int act, float slope, int flip, __nv_bfloat16* y_bf16, float* y_nchw, cudaStream_t s) {
  const int R = thin_out_rows(c);
  const int smem = (8 * (9 * 64 + 8) + (R + 2) * (c.W + 2) * kLdC) * 2;
  static int configured = 0;
  if (int rc = set_smem(thin_out_mma_kernel, smem, &configured)) return rc;
  dim3 grid((c.H + R - 1) / R, c.N);
  thin_out_mma_kernel<<<grid, kThreads, smem, s>>>(c, R, x, w, bias, act, slope, flip, y_bf16, y_nchw);
  return check("thin_out_conv");
}

int thin_wgrad(const ThinConv& c, const __nv_bfloat16* small, const __nv_bfloat16* wide, int sgn,
               float* out, float* small_colsum, float* wide_colsum, cudaStream_t s) {
  const int T = c.k * c.k;
  cudaMemsetAsync(out, 0, sizeof(float) * T * c.CS * c.CW, s);
  if (wide_colsum) cudaMemsetAsync(wide_colsum, 0, sizeof(float) * c.CW, s);
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  const int ntiles = static_cast<int>((npix + kTileQ - 1) / kTileQ);
  if (c.k == 3) {
    const int smem = (kTileQ * kLdC + 32 * (kTileQ + 8)) * 2;
    dim3 grid(ntiles < 148 * 4 ? ntiles : 148 * 4, c.CW / 64);
    thin_wgrad_mma_kernel<1><<<grid, kThreads, smem, s>>>(c, small, wide, sgn, out, wide_colsum, ntiles);
  } else {
    const int smem = (kTileQ * kLdC + 256 * (kTileQ + 8)) * 2;
    static int configured = 0;
    if (int rc = set_smem(thin_wgrad_mma_kernel<8>, smem, &configured)) return rc;
    dim3 grid(ntiles < 148 ? ntiles : 148, c.CW / 64);
    thin_wgrad_mma_kernel<8><<<grid, kThreads, smem, s>>>(c, small, wide, sgn, out, wide_colsum, ntiles);
  }
  if (small_colsum) {
    cudaMemsetAsync(small_colsum, 0, sizeof(float) * c.CS, s);
    thin_colsum_kernel<3><<<148, 256, 0, s>>>(small, npix, small_colsum);
  }
  return check("thin_wgrad");
}

}  // namespace sisr
