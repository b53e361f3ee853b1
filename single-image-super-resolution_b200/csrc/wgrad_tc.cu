// tcgen05 weight-gradient engine for 3x3 convolutions (stride 1/2, pad 1) on NHWC bf16.
//
//   G[co, tap, ci] = sum over output pixels p of  dY[p, co] * X[p shifted by tap, ci]
//
// The reduction dimension is the pixel index, so both UMMA operands are MN-major: the shared
// memory tiles are [64 pixels][64 channels] (128 B rows, hardware 128 B swizzle), exactly what the
// im2col-mode TMA of the forward pass delivers.
//   A operand (M = 128): two im2col tiles of X for two filter taps (64 ci each); the second
//                        64-row group is reached through the descriptor's leading-dimension offset
//   B operand (N = 64) : one tile of dY (64 output channels)
//   accumulators       : 5 tap-pairs x 64 fp32 columns in TMEM (tap 9 of the last pair is unused)
// One CTA owns (ci block, co block, pixel range); pixel ranges are split across CTAs and the fp32
// partials are summed by a second kernel (deterministic, no atomics).
#include "wgrad_tc.h"

#include <stdio.h>
#include <stdlib.h>

#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace sisr {

namespace {

constexpr int kPix = 64;                      // pixels per k-block
constexpr int kTile = kPix * 64 * 2;          // 8 KB
constexpr int kSlots = 11;                    // 9 taps + 1 unused (pair of tap 8) + dY
constexpr int kStageBytes = kSlots * kTile;   // 88 KB
constexpr int kStages = 2;                    // (32-pixel k-blocks x 4 stages and a halo-box feed measured no
                                              // faster, profiles/r1_notes.md: bound by the N = 64 instruction floor)
constexpr int kThreads = 320;                 // TMA warp, MMA warp, 8 epilogue warps (quadrant x column half)
constexpr int kPairs = 5;
constexpr int kTmemCols = 512;

struct WParams {
  int M;                 // output pixels N*OH*OW
  int OH, OW, stride;
  int ci_blocks, co_blocks, splits, kb_per_split, total_kb;
  int Cin, Cout;
  int ps;                // dY stored pixel-shuffled
  int atomic;            // 1: every split adds into ONE zeroed [Cout][9][Cin] buffer (red.global.add.f32)
  float* out;            // [splits][Cout][9][Cin] (splits > 1) or final [Cout][9][Cin]
  long long* dbg;        // nullable (tools/wgrad_phases.py): phase cycle counters of CTA 0, see wgrad_tc_set_debug
};

thread_local char g_err[256] = "";
long long* g_wgrad_dbg = nullptr;

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                const WParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int split = b % p.splits;
  b /= p.splits;
  const int cob = b % p.co_blocks;
  const int cib = b / p.co_blocks;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.total_kb, kb0 + p.kb_per_split);
  const int num_kb = kb1 - kb0;
  const long long k0 = p.dbg ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int hw = p.OH * p.OW;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kStages;
        const uint32_t round = i / kStages;
        const long long w0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&empty_bar[s]), (round & 1) ^ 1);
        if (p.dbg && blockIdx.x == 0) p.dbg[0] += clock64() - w0;      // producer waits for a free stage
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, 10 * kTile);
        const int p0 = (kb0 + i) * kPix;
        const int n_img = p0 / hw;
        const int rem = p0 - n_img * hw;
        const int gh = rem / p.OW;
        const int gw = rem - gh * p.OW;
        const uint32_t base = smem_u32(smem + s * kStageBytes);
        for (int tap = 0; tap < 9; ++tap)
          tma_load_im2col_4d(base + tap * kTile, &tmap_x, fb, cib * 64, gw * p.stride - 1,
                             gh * p.stride - 1, n_img, static_cast<uint16_t>(tap % 3),
                             static_cast<uint16_t>(tap / 3));
        if (p.ps)
          tma_load_im2col_4d(base + 10 * kTile, &tmap_dy, fb, 0, gw * 2, gh * 2, n_img,
                             static_cast<uint16_t>(cob & 1), static_cast<uint16_t>(cob >> 1));
        else
          tma_load_im2col_4d(base + 10 * kTile, &tmap_dy, fb, cob * 64, gw, gh, n_img, 0, 0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % kStages;
      const uint32_t round = i / kStages;
      const long long c0 = p.dbg ? clock64() : 0;
      mbar_wait(smem_u32(&full_bar[s]), round & 1);
      const long long c1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      if (lane == 0) {
        const uint32_t base = smem_u32(smem + s * kStageBytes);
        const uint32_t dy_addr = base + 10 * kTile;
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
#pragma unroll
          for (int j = 0; j < kPix / 16; ++j) {
            // MN-major, 128 B swizzle: LBO = distance between the two 64-channel groups (= one
            // tile), SBO = 8 pixel rows = 1024 B; a K step of 16 pixels advances 2048 B.
            const uint64_t da = umma_smem_desc(base + (2 * q) * kTile + j * 2048, kTile, 1024);
            const uint64_t db = umma_smem_desc(dy_addr + j * 2048, kTile, 1024);
            umma_bf16(tmem_base + q * 64, da, db, idesc, (i > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        if (i == num_kb - 1) umma_commit(smem_u32(&tmem_full_bar));
        if (p.dbg && blockIdx.x == 0) {
          p.dbg[1] += c1 - c0;              // MMA warp waits for the operands of a k-block
          p.dbg[2] += clock64() - c1;       // ... issues its 20 instructions
          p.dbg[3] += 1;
        }
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;          // which 32 of the 64 output channels of every tap pair
    const int row = quad * 32 + lane;          // accumulator row: tap parity * 64 + ci
    const int ci = cib * 64 + (row & 63);
    mbar_wait(smem_u32(&tmem_full_bar), 0);
    const long long e0 = p.dbg ? clock64() : 0;
    tc_fence_after();
    float* outp = p.out + (p.atomic ? 0 : static_cast<size_t>(split) * p.Cout * 9 * p.Cin);
#pragma unroll 1
    for (int q = 0; q < kPairs; ++q) {
      const int tap = 2 * q + (row >> 6);
      {
        const int c = half;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + q * 64 + c * 32, raw);
        tmem_ld_wait();
        if (tap < 9) {
          if (p.atomic) {
            // small layers (the 64 x 576 trunk gradient): the pixel splits add straight into one L2-resident
            // buffer (a warp's 32 lanes = 32 consecutive ci = one 128-byte reduction) instead of writing
            // `splits` partial copies that a second kernel reads back
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = cob * 64 + c * 32 + i;
              atomicAdd(&outp[(static_cast<size_t>(co) * 9 + tap) * p.Cin + ci], __uint_as_float(raw[i]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = cob * 64 + c * 32 + i;
              outp[(static_cast<size_t>(co) * 9 + tap) * p.Cin + ci] = __uint_as_float(raw[i]);
            }
          }
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) p.dbg[4] += clock64() - e0;   // epilogue (stores issued)
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    p.dbg[5] += clock64() - k0;
    p.dbg[6] += 1;
    p.dbg[7] = gridDim.x;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// out[i] = sum_k ws[k][i]: block = 16 float4 columns x 16 split lanes, so that a thread issues at
// most ceil(splits/16) independent 16-byte loads (the partials are 20-150 MB in total: HBM-bound).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, long long total4,
                     int splits) {
  __shared__ float4 s_part[16][16];
  const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
  const long long i = static_cast<long long>(blockIdx.x) * 16 + x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < total4) {
    const float4* p = reinterpret_cast<const float4*>(ws) + i;
#pragma unroll 4
    for (int k = y; k < splits; k += 16) {
      const float4 v = __ldg(p + static_cast<long long>(k) * total4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  s_part[y][x] = acc;
  __syncthreads();
  if (y == 0 && i < total4) {
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const float4 v = s_part[k][x];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// dbias[cls * C + c] = sum over rows of class cls of dy[row, c]; classes = PixelShuffle sub-pixels
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, long long rows, int C, int W2,
                                 int ps, float* __restrict__ dbias) {
  extern __shared__ float s_acc[];   // [classes * C]
  const int classes = ps ? 4 : 1;
  for (int i = threadIdx.x; i < classes * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int tpr = C / 8;
  const int rpb = blockDim.x / tpr;          // even (C <= 1024), so a thread's rows keep their x parity
  const int r = threadIdx.x / tpr, cv = threadIdx.x % tpr;
  if (r < rpb) {
    float acc[2][8];                         // [y parity][channel]; without PixelShuffle only [0]
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
    const long long first = static_cast<long long>(blockIdx.x) * rpb + r;
    for (long long row = first; row < rows; row += static_cast<long long>(gridDim.x) * rpb) {
      const bool odd = ps && ((row / W2) & 1);
      const uint4 u = *reinterpret_cast<const uint4*>(dy + row * C + cv * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        acc[0][2 * j] += odd ? 0.f : f.x;
        acc[0][2 * j + 1] += odd ? 0.f : f.y;
        acc[1][2 * j] += odd ? f.x : 0.f;
        acc[1][2 * j + 1] += odd ? f.y : 0.f;
      }
    }
    const int xpar = ps ? static_cast<int>(first & 1) : 0;
#pragma unroll
    for (int yp = 0; yp < 2; ++yp) {
      if (yp && !ps) break;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(yp * 2 + xpar) * C + cv * 8 + j], acc[yp][j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < classes * C; i += blockDim.x) atomicAdd(&dbias[i], s_acc[i]);
}

// gradients up to this many floats accumulate with L2 reductions: every layer of the step (the largest, 512 x 512
// x 9, is 9.4 MB and stays L2-resident).  1 MB (trunk only) -> all: step 8.09 -> 7.95 ms sustained;
// SISR_WGRAD_ATOMIC_MAX overrides for A-B timing
const long long kAtomicMaxFloats = [] { const char* e = getenv("SISR_WGRAD_ATOMIC_MAX"); return e ? atoll(e) : (1ll << 22); }();
int g_wgrad_atomic = [] { const char* e = getenv("SISR_WGRAD_ATOMIC"); return e && e[0] == '0' ? 0 : 1; }();

struct Plan {
  int total_kb, splits, kb_per_split, tiles, atomic;
};
Plan make_plan(int n, int oh, int ow, int cin, int cout, int stride = 0, int ps = 1) {
  Plan pl{};
  const long long M = static_cast<long long>(n) * oh * ow;
  pl.total_kb = static_cast<int>((M + kPix - 1) / kPix);
  pl.tiles = (cin / 64) * (cout / 64);
  double fixed_cost = 1.0;       // per-CTA prologue + partial write, in k-blocks (a k-block = 20 MMAs ~ 1.4 us)
  // split the pixel range so that tiles x splits fills whole waves of the 148 SMs: minimise
  // waves / splits (time of the slowest SM), ties go to fewer splits (less partial-sum traffic)
  int splits = 1;
  double best = 1e30;
  // cost of one more split, in k-blocks: its fp32 partial (Cout x 9 x Cin) is written once and read once by the
  // finish kernel, and the CTA pays its prologue / pipeline fill.  SISR_WGRAD_SPLIT_COST overrides (A-B timing).
  static const double split_cost = [] {
    const char* e = getenv("SISR_WGRAD_SPLIT_COST");
    return e ? atof(e) : 0.1;      // 0.02 -> 0.1: step 8.38 -> 8.21 ms (r2 sweep 0.02 / 0.1 / 0.2 / 0.4)
  }();
  pl.atomic = (g_wgrad_atomic && static_cast<long long>(cout) * 9 * cin <= kAtomicMaxFloats) ? 1 : 0;
  const double per_split = pl.atomic ? 0.03 : split_cost;     // no partial copy to write and read back
  const int max_splits = pl.total_kb < 148 ? pl.total_kb : 148;
  for (int sp = 1; sp <= max_splits; ++sp) {
    const int kbs = (pl.total_kb + sp - 1) / sp;
    const int real = (pl.total_kb + kbs - 1) / kbs;
    const long long waves = (static_cast<long long>(pl.tiles) * real + 147) / 148;
    // the kernel is MMA-bound (20 N=64 instructions per k-block): spread over all SMs, avoid partial waves
    const double cost = static_cast<double>(waves) * (kbs + fixed_cost) + per_split * real;
    if (cost < best - 1e-9) {
      best = cost;
      splits = real;
    }
  }
  pl.kb_per_split = (pl.total_kb + splits - 1) / splits;
  pl.splits = (pl.total_kb + pl.kb_per_split - 1) / pl.kb_per_split;
  if (pl.splits <= 1) pl.atomic = 0;
  return pl;
}

}  // namespace

const char* wgrad_tc_last_error() { return g_err; }
void wgrad_tc_set_debug(long long* counters) { g_wgrad_dbg = counters; }

bool wgrad_tc_supported(int n, int h, int w, int cin, int oh, int ow, int cout, int k, int stride,
                        int pad, int ps_r) {
  if (k != 3 || pad != 1 || (stride != 1 && stride != 2)) return false;
  if (cin % 64 || cout % 64) return false;
  if (ps_r == 2 && (stride != 1 || cout / 4 != 64)) return false;
  return n > 0 && h > 0 && w > 0;
}

size_t wgrad_tc_workspace_bytes(int n, int h, int w, int cin, int oh, int ow, int cout, int k,
                                int stride, int pad, int ps_r) {
  if (!wgrad_tc_supported(n, h, w, cin, oh, ow, cout, k, stride, pad, ps_r)) return 0;
  const Plan pl = make_plan(n, oh, ow, cin, cout, stride, ps_r == 2);
  return static_cast<size_t>(pl.splits) * cout * 9 * cin * sizeof(float);
}

int wgrad_tc_launch(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* g, float* dbias,
                    void* workspace, int n, int h, int w, int cin, int oh, int ow, int cout, int stride,
                    int ps_r, cudaStream_t s, int* splits_out) {
  Plan pl = make_plan(n, oh, ow, cin, cout, stride, ps_r == 2);
  const bool keep_partials = g == nullptr;   // caller reduces (weight_grad_reduce_finish)
  if ((pl.splits > 1 || keep_partials) && !workspace) {
    snprintf(g_err, sizeof g_err, "wgrad: split-K workspace missing");
    return 1;
  }
  if (pl.atomic)     // the splits add into the first [Cout][9][Cin] slot of the workspace
    cudaMemsetAsync(workspace, 0, sizeof(float) * static_cast<size_t>(cout) * 9 * cin, s);
  if (splits_out) *splits_out = pl.atomic ? 1 : pl.splits;
  CUtensorMap tx, tdy;
  if (make_tmap_im2col_nhwc_bf16(&tx, x, n, h, w, cin, -1, -1, -1, -1, 64, kPix, stride)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  const int ps = ps_r == 2;
  int rc;
  if (ps)
    rc = make_tmap_im2col_nhwc_bf16(&tdy, dy, n, 2 * oh, 2 * ow, cout / 4, 0, 0, -1, -1, 64, kPix, 2);
  else
    rc = make_tmap_im2col_nhwc_bf16(&tdy, dy, n, oh, ow, cout, 0, 0, 0, 0, 64, kPix, 1);
  if (rc) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  WParams p;
  p.M = n * oh * ow;
  p.OH = oh; p.OW = ow; p.stride = stride;
  p.ci_blocks = cin / 64; p.co_blocks = cout / 64;
  p.splits = pl.splits; p.kb_per_split = pl.kb_per_split; p.total_kb = pl.total_kb;
  p.Cin = cin; p.Cout = cout; p.ps = ps;
  p.atomic = pl.atomic;
  p.dbg = g_wgrad_dbg;
  p.out = (pl.splits > 1 || keep_partials) ? static_cast<float*>(workspace) : g;
  const int smem_bytes = kStages * kStageBytes + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 3;
    }
    configured = true;
  }
  const int grid = pl.tiles * pl.splits;
  wgrad_tc_kernel<<<grid, kThreads, smem_bytes, s>>>(tx, tdy, p);
  if (pl.splits > 1 && !keep_partials) {
    const long long total4 = static_cast<long long>(cout) * 9 * cin / 4;
    splitk_reduce_kernel<<<static_cast<int>((total4 + 15) / 16), 256, 0, s>>>(
        static_cast<const float*>(workspace), g, total4, pl.atomic ? 1 : pl.splits);
  }
  if (dbias) {
    cudaMemsetAsync(dbias, 0, sizeof(float) * cout, s);
    const int C = ps ? cout / 4 : cout;
    const long long rows = ps ? static_cast<long long>(n) * 4 * oh * ow : static_cast<long long>(n) * oh * ow;
    if (C % 8 || C / 8 > 256) {
      snprintf(g_err, sizeof g_err, "wgrad: bias gradient needs C %% 8 == 0, C <= 2048");
      return 1;
    }
    const int rpb = 256 / (C / 8);
    long long blocks = (rows + rpb - 1) / rpb;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (ps && ((blocks * rpb) & 1)) ++blocks;      // a thread's rows must keep their x parity (even row stride)
    bias_grad_kernel<<<static_cast<int>(blocks), 256, (ps ? 4 : 1) * C * sizeof(float), s>>>(
        dy, rows, C, 2 * ow, ps, dbias);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "wgrad launch: %s", cudaGetErrorString(e));
    return 4;
  }
  return 0;
}

}  // namespace sisr
