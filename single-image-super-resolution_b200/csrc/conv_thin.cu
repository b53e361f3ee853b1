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
#include <stdlib.h>

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

// ------------------------------------------------------------------ streaming 3x3 kernels (v2)
// The first-generation kernels above stage an im2col tile per 128 pixels behind block-wide barriers
// and run 5-6x off the HBM time of the 75 MB tensor they stream (profiles/r1_notes.md).  The kernels
// below keep the mma.sync contraction but remove the staging: persistent warps, weights held as B
// fragments in registers for the whole kernel, A fragments assembled straight from global memory.

// n / d for n < 2^31 (Granlund-Montgomery): q = (n * m) >> (31 + l), m = ceil(2^(31+l) / d)
struct FastDiv {
  uint32_t m, sh, d;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return static_cast<uint32_t>((static_cast<unsigned long long>(n) * f.m) >> f.sh);
}
__device__ __forceinline__ uint32_t ldg_u16(const __nv_bfloat16* p) {
  return static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p)));
}

// thin-in, 3 -> 64, 3x3, pad 1.  One warp = 16 consecutive pixels x 64 output channels per trip:
// K = 27 (+5 zero) in two k-steps; the eight k indices a thread feeds (m16n8k16 A layout) are fixed
// for the whole kernel, so their tap offsets are constants and a pixel costs one 9-bit validity mask.
// Output rows leave through a per-warp 2 KB staging tile (no block barrier) as 16-byte NHWC stores.
__global__ void __launch_bounds__(kThreads, 2)
thin_in3_kernel(ThinConv c, FastDiv dW, FastDiv dH, int npix, const __nv_bfloat16* __restrict__ x,
                const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias, int act, float slope,
                const float* __restrict__ slope_ptr, int flip, __nv_bfloat16* __restrict__ y) {
  __shared__ __align__(16) __nv_bfloat16 stage[kThreads / 32][16 * kLdC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int cwb = blockIdx.y * 64;
  const unsigned short* wu = reinterpret_cast<const unsigned short*>(w);

  uint32_t bfr[2][8][2];
  float bb[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int n = cwb + nt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = ks * 16 + r * 8 + 2 * t + e;
          if (k < 27) {
            const int tap = k / 3, cs = k - tap * 3;
            v |= static_cast<uint32_t>(__ldg(wu + (static_cast<size_t>(n) * 9 + (flip ? 8 - tap : tap)) * 3 + cs))
                 << (16 * e);
          }
        }
        bfr[ks][nt][r] = v;
      }
    bb[nt][0] = bias ? bias[cwb + nt * 8 + 2 * t] : 0.f;
    bb[nt][1] = bias ? bias[cwb + nt * 8 + 2 * t + 1] : 0.f;
  }
  // the thread's eight k indices: i = ks*4 + r*2 + e  <->  k = ks*16 + r*8 + 2t + e
  int off[8], tapi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = (i >> 2) * 16 + ((i >> 1) & 1) * 8 + 2 * t + (i & 1);
    const int tap = k / 3, cs = k - tap * 3;
    const int kh = tap / 3, kw = tap - kh * 3;
    tapi[i] = tap;                                   // 9, 10 for the zero columns: mask bits are 0 there
    off[i] = k < 27 ? ((kh - 1) * c.W + (kw - 1)) * 3 + cs : 0;
  }
  if (act == ACT_PRELU) slope = *slope_ptr;
  if (act == ACT_RELU) slope = 0.f;
  if (act == ACT_NONE) slope = 1.f;                  // x > 0 ? x : x * 1

  const int ntiles = (npix + 15) >> 4;
  const int stride = gridDim.x * (kThreads / 32);

  auto load_a = [&](int mt, uint32_t (&a)[8]) {
    uint32_t h[2][8];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int q = mt * 16 + g + p * 8;
      const bool valid = q < npix;
      const uint32_t qq = valid ? q : 0;
      const uint32_t t1 = fdiv(qq, dW);
      const int ow = qq - t1 * c.W;
      const int oh = t1 - fdiv(t1, dH) * c.H;
      const uint32_t cb = (ow > 0 ? 1u : 0u) | 2u | (ow < c.W - 1 ? 4u : 0u);
      uint32_t mask = (oh > 0 ? cb : 0u) | (cb << 3) | (oh < c.H - 1 ? cb << 6 : 0u);
      if (!valid) mask = 0;
      const __nv_bfloat16* xp = x + static_cast<size_t>(qq) * 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) h[p][i] = ((mask >> tapi[i]) & 1u) ? ldg_u16(xp + off[i]) : 0u;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      a[ks * 4 + 0] = h[0][ks * 4 + 0] | (h[0][ks * 4 + 1] << 16);    // row g,   k = 2t, 2t+1
      a[ks * 4 + 1] = h[1][ks * 4 + 0] | (h[1][ks * 4 + 1] << 16);    // row g+8
      a[ks * 4 + 2] = h[0][ks * 4 + 2] | (h[0][ks * 4 + 3] << 16);    // row g,   k + 8
      a[ks * 4 + 3] = h[1][ks * 4 + 2] | (h[1][ks * 4 + 3] << 16);    // row g+8
    }
  };

  int mt = blockIdx.x * (kThreads / 32) + warp;
  uint32_t a[8], an[8];
  if (mt < ntiles) load_a(mt, a);
  __nv_bfloat16* st = stage[warp];
  while (mt < ntiles) {
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        mma_bf16(acc[nt], a[ks * 4], a[ks * 4 + 1], a[ks * 4 + 2], a[ks * 4 + 3], bfr[ks][nt][0], bfr[ks][nt][1]);
    const int mn = mt + stride;
    if (mn < ntiles) load_a(mn, an);                 // in flight during the epilogue
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      float v0 = acc[nt][0] + bb[nt][0], v1 = acc[nt][1] + bb[nt][1];
      float v2 = acc[nt][2] + bb[nt][0], v3 = acc[nt][3] + bb[nt][1];
      v0 = v0 > 0.f ? v0 : v0 * slope; v1 = v1 > 0.f ? v1 : v1 * slope;
      v2 = v2 > 0.f ? v2 : v2 * slope; v3 = v3 > 0.f ? v3 : v3 * slope;
      *reinterpret_cast<uint32_t*>(st + g * kLdC + col) = pack_bf16x2(v0, v1);
      *reinterpret_cast<uint32_t*>(st + (g + 8) * kLdC + col) = pack_bf16x2(v2, v3);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 4 + (lane >> 3), chunk = lane & 7;
      const int q = mt * 16 + row;
      if (q < npix)
        *reinterpret_cast<uint4*>(y + static_cast<size_t>(q) * c.CW + cwb + chunk * 8) =
            *reinterpret_cast<const uint4*>(st + row * kLdC + chunk * 8);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = an[i];
    mt = mn;
  }
}

// thin-out, 64 -> 3, 3x3, pad 1, as GEMM + col2im.  Phase A: for every input pixel of the block's
// (R+2)-row band P[pixel][tap*3+cs] = sum_c x[pixel][c] * w[cs][tap][c]  (one [16 x 64] x [64 x 32]
// product per 16 pixels: 16 mma.sync instead of the 36 of a gather formulation, and every input
// element is read exactly once, as two full 64-byte half-rows per thread quad -- the contraction
// order over the 64 channels is permuted so that a thread's A fragment IS its two 16-byte loads).
// P stays in shared memory (fp32, row stride 29 words: conflict-free for the gather).  Phase B: every
// output pixel adds the nine P entries of its neighbours, + bias, Tanh / identity, bf16 NHWC and/or
// fp32 NCHW store.
constexpr int kLdP = 29;
__global__ void __launch_bounds__(kThreads, 2)
thin_out3_kernel(ThinConv c, int R, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                 const float* __restrict__ bias, int act, float slope, int flip,
                 __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_nchw) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* P = reinterpret_cast<float*>(smem_raw);            // [(R+2)*W][kLdP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n = blockIdx.y, row0 = blockIdx.x * R;
  const unsigned short* wu = reinterpret_cast<const unsigned short*>(w);

  // B fragments: k-step ks, register r, element e  <->  physical channel (ks>>1)*32 + 8t + (ks&1)*4 + r*2 + e
  uint32_t bfr[4][4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int j = nt * 8 + g;
    const int tap = j / 3, cs = j - tap * 3;
    const size_t wrow = (static_cast<size_t>(cs) * 9 + (flip ? 8 - tap : tap)) * 64;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ch = (ks >> 1) * 32 + 8 * t + (ks & 1) * 4 + r * 2;
        bfr[ks][nt][r] = j < 27 ? *reinterpret_cast<const uint32_t*>(wu + wrow + ch) : 0u;
      }
  }

  const int E = (R + 2) * c.W;
  const long long img0 = static_cast<long long>(n) * c.H * c.W;
  const long long lin0 = img0 + static_cast<long long>(row0 - 1) * c.W;   // linear pixel of band entry 0
  const long long img1 = img0 + static_cast<long long>(c.H) * c.W;
  const int ntiles = (E + 15) >> 4;

  auto load_a = [&](int mt, uint4 (&v)[4]) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int e = mt * 16 + g + p * 8;
      const long long lin = lin0 + e;
      const bool ok = e < E && lin >= img0 && lin < img1;
      const uint4* src = reinterpret_cast<const uint4*>(x + (ok ? lin : img0) * 64 + 8 * t);
      v[p * 2] = ok ? __ldg(src) : make_uint4(0, 0, 0, 0);
      v[p * 2 + 1] = ok ? __ldg(src + 4) : make_uint4(0, 0, 0, 0);
    }
  };

  int mt = warp;
  uint4 v[4], vn[4];
  if (mt < ntiles) load_a(mt, v);
  while (mt < ntiles) {
    const int mn = mt + kThreads / 32;
    if (mn < ntiles) load_a(mn, vn);
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint4& lo = v[ks >> 1];          // row g
      const uint4& hi = v[2 + (ks >> 1)];    // row g + 8
      const uint32_t a0 = (ks & 1) ? lo.z : lo.x, a2 = (ks & 1) ? lo.w : lo.y;
      const uint32_t a1 = (ks & 1) ? hi.z : hi.x, a3 = (ks & 1) ? hi.w : hi.y;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], a0, a1, a2, a3, bfr[ks][nt][0], bfr[ks][nt][1]);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int e = mt * 16 + g + p * 8;
      if (e < E) {
        float* row = P + e * kLdP;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = nt * 8 + 2 * t;
          if (col < 27) row[col] = acc[nt][p * 2];
          if (col + 1 < 27) row[col + 1] = acc[nt][p * 2 + 1];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = vn[i];
    mt = mn;
  }
  __syncthreads();

  const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f, b2 = bias ? bias[2] : 0.f;
  const int npx = R * c.W;
  for (int p = tid; p < npx; p += kThreads) {
    const int r = p / c.W, ow = p - r * c.W, oh = row0 + r;
    if (oh >= c.H) break;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh + kh - 1;
      if (ih < 0 || ih >= c.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow + kw - 1;
        if (iw < 0 || iw >= c.W) continue;
        const float* e = P + ((r + kh) * c.W + iw) * kLdP + (kh * 3 + kw) * 3;
        s0 += e[0]; s1 += e[1]; s2 += e[2];
      }
    }
    s0 = act_apply(s0 + b0, act, slope);
    s1 = act_apply(s1 + b1, act, slope);
    s2 = act_apply(s2 + b2, act, slope);
    const size_t q = static_cast<size_t>(img0) + static_cast<size_t>(oh) * c.W + ow;
    if (y_bf16) {
      y_bf16[q * 3] = __float2bfloat16_rn(s0);
      y_bf16[q * 3 + 1] = __float2bfloat16_rn(s1);
      y_bf16[q * 3 + 2] = __float2bfloat16_rn(s2);
    }
    if (y_nchw) {
      const size_t hw = static_cast<size_t>(c.H) * c.W;
      float* o = y_nchw + static_cast<size_t>(n) * 3 * hw + static_cast<size_t>(oh) * c.W + ow;
      o[0] = s0; o[hw] = s1; o[2 * hw] = s2;
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

// thin weight gradient, 3x3, pad 1 (same contraction as thin_wgrad_mma_kernel<1>, streaming form):
// one warp = 16 pixels per trip.  A = wide^T from a per-warp double-buffered cp.async tile (ldmatrix.trans),
// B = the im2col of `small` assembled straight from global memory: the four columns j = nt*8 + g a thread
// feeds are fixed, so tap offsets are constants; column 27 is all ones (bias gradient of a thin-in conv).
// No block barrier inside the loop; 64 fp32 accumulators per thread live across the whole kernel.
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kThreads, 2)
thin_wgrad3_kernel(ThinConv c, FastDiv dW, FastDiv dH, int npix, const __nv_bfloat16* __restrict__ small,
                   const __nv_bfloat16* __restrict__ wide, int sgn, float* __restrict__ out,
                   float* __restrict__ wide_colsum) {
  __shared__ __align__(16) __nv_bfloat16 stage[kThreads / 32][2][16 * kLdC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int cwb = blockIdx.y * 64;
  int off[4];
  uint32_t bit[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int j = nt * 8 + g;
    const int tap = j / 3, cs = j - tap * 3, kh = tap / 3, kw = tap - kh * 3;
    off[nt] = j < 27 ? sgn * ((kh - 1) * c.W + (kw - 1)) * 3 + cs : 0;
    bit[nt] = j < 27 ? 1u << tap : (j == 27 ? 1u << 9 : 0u);     // bit 9 of a pixel mask = "pixel exists"
  }
  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  const int ntiles = (npix + 15) >> 4;
  const int stride = gridDim.x * (kThreads / 32);
  auto stage_tile = [&](int mt, int buf) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 4 + (lane >> 3), chunk = lane & 7;
      __nv_bfloat16* dst = stage[warp][buf] + row * kLdC + chunk * 8;
      const int q = mt * 16 + row;
      if (q < npix)
        cp_async16(smem_u32(dst), wide + static_cast<size_t>(q) * c.CW + cwb + chunk * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
  };
  // validity mask of a pixel: bit tap = the tap's source pixel is inside the image; bit 9 = pixel exists
  auto pixel_mask = [&](int q) -> uint32_t {
    if (q >= npix) return 0u;
    const uint32_t t1 = fdiv(q, dW);
    const int qw = q - t1 * c.W;
    const int qh = t1 - fdiv(t1, dH) * c.H;
    const uint32_t wl = qw > 0, wh = qw < c.W - 1, hl = qh > 0, hh = qh < c.H - 1;
    const uint32_t cb = sgn > 0 ? (wl | 2u | (wh << 2)) : (wh | 2u | (wl << 2));
    const uint32_t r0 = sgn > 0 ? hl : hh, r2 = sgn > 0 ? hh : hl;
    return (r0 ? cb : 0u) | (cb << 3) | (r2 ? cb << 6 : 0u) | (1u << 9);
  };

  int mt = blockIdx.x * (kThreads / 32) + warp;
  int buf = 0;
  if (mt < ntiles) stage_tile(mt, 0);
  while (mt < ntiles) {
    const int mn = mt + stride;
    if (mn < ntiles) stage_tile(mn, buf ^ 1); else cp_async_commit();
    // B fragments of this tile
    uint32_t b[4][2];
    {
      uint32_t h[4][4];
#pragma unroll
      for (int pi = 0; pi < 4; ++pi) {
        const int q = mt * 16 + 2 * t + (pi & 1) + (pi >> 1) * 8;
        const uint32_t mask = pixel_mask(q);
        const __nv_bfloat16* sp = small + static_cast<size_t>(q < npix ? q : 0) * 3;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t v = (mask & bit[nt]) ? ldg_u16(sp + off[nt]) : 0u;
          if (nt == 3 && g == 3) v = (mask >> 9) ? 0x3F80u : 0u;       // j = 27: ones column
          h[nt][pi] = v;
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b[nt][0] = h[nt][0] | (h[nt][1] << 16);
        b[nt][1] = h[nt][2] | (h[nt][3] << 16);
      }
    }
    cp_async_wait<1>();
    __syncwarp();
    const __nv_bfloat16* Wd = stage[warp][buf];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4_trans(smem_u32(Wd + ((lane & 7) + ((lane >> 4) << 3)) * kLdC + m * 16 + ((lane >> 3) & 1) * 8),
                    a0, a1, a2, a3);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[m][nt], a0, a1, a2, a3, b[nt][0], b[nt][1]);
    }
    __syncwarp();
    buf ^= 1;
    mt = mn;
  }
  cp_async_wait<0>();
  __syncthreads();
  // combine the 8 warps, then one atomic per (cw, j) and block
  float* red = reinterpret_cast<float*>(&stage[0][0][0]);    // [64][33]
  for (int wv = 0; wv < kThreads / 32; ++wv) {
    if (warp == wv) {
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cw = m * 16 + g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
            float* r = red + cw * 33 + j;
            *r = (wv == 0) ? acc[m][nt][e] : *r + acc[m][nt][e];
          }
    }
    __syncthreads();
  }
  for (int i = tid; i < 64 * 28; i += kThreads) {
    const int cw = i / 28, j = i - cw * 28;
    const float v = red[cw * 33 + j];
    if (j < 27) {
      const int tap = j / 3, cs = j - tap * 3;
      float* dst = (sgn > 0) ? out + static_cast<size_t>(cwb + cw) * 27 + j
                             : out + (static_cast<size_t>(cs) * 9 + tap) * c.CW + cwb + cw;
      atomicAdd(dst, v);
    } else if (wide_colsum) {
      atomicAdd(&wide_colsum[cwb + cw], v);
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

FastDiv make_fastdiv(uint32_t d) {
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned long long one = 1ull << (31 + l);
  return FastDiv{static_cast<uint32_t>((one + d - 1) / d), 31 + l, d};
}
// (the first-generation thin_*_mma kernels below still serve the 9x9 layer; their use for the 3x3 layers
// was an A-B switch in round 1 and is gone)
constexpr bool thin_v1() { return false; }
// rows per block of thin_out3: the fp32 tap-product band (R+2) x W x 29 words must fit; two blocks per SM
// when that leaves at least 4 rows, else one
int thin_out3_rows(const ThinConv& c) {
  const int per_row = c.W * kLdP * 4;
  int r = 111 * 1024 / per_row - 2;
  if (r < 4) r = 222 * 1024 / per_row - 2;
  if (r > c.H) r = c.H;
  if (r > 16) r = 16;
  return r;      // < 1: not supported
}

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
  if (c.k == 3 && c.pad == 1 && act != ACT_TANH && npix < (1ll << 31) - 16 && !thin_v1()) {
    const int ntiles = static_cast<int>((npix + 15) >> 4);
    const int blocks = (ntiles + 7) / 8;
    dim3 grid3(blocks < 148 * 2 ? blocks : 148 * 2, c.CW / 64);
    thin_in3_kernel<<<grid3, kThreads, 0, s>>>(c, make_fastdiv(c.W), make_fastdiv(c.H), static_cast<int>(npix), x,
                                               w, bias, act, slope, slope_ptr, flip, y);
    return check("thin_in3");
  }
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
  if (const int R3 = thin_out3_rows(c); R3 >= 1 && c.pad == 1 && !thin_v1()) {
    const int smem3 = (R3 + 2) * c.W * kLdP * 4;
    static int configured3 = 0;
    if (int rc = set_smem(thin_out3_kernel, smem3, &configured3)) return rc;
    dim3 grid3((c.H + R3 - 1) / R3, c.N);
    thin_out3_kernel<<<grid3, kThreads, smem3, s>>>(c, R3, x, w, bias, act, slope, flip, y_bf16, y_nchw);
    return check("thin_out3");
  }
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
  if (c.k == 3 && c.pad == 1 && npix < (1ll << 31) - 16 && !thin_v1()) {
    const int groups = static_cast<int>((npix + 15) >> 4);
    const int blocks = (groups + 7) / 8;
    dim3 grid(blocks < 148 * 2 ? blocks : 148 * 2, c.CW / 64);
    thin_wgrad3_kernel<<<grid, kThreads, 0, s>>>(c, make_fastdiv(c.W), make_fastdiv(c.H), static_cast<int>(npix),
                                                 small, wide, sgn, out, wide_colsum);
  } else if (c.k == 3) {
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
