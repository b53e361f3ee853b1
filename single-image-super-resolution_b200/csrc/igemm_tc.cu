// tcgen05 implicit-GEMM convolution engine for sm_100a.
//
//   warp 0      : TMA producer
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  : epilogue (tcgen05.ld -> bias / activation / BN statistics -> bf16 NHWC store)
//
// Persistent: one CTA per SM walks the 128 x BN output tiles (tile = blockIdx.x + i * gridDim.x).
// The TMEM accumulator is double buffered (2 x BN columns), so the epilogue of tile i overlaps the
// main loop of tile i+1 and barrier / TMEM set-up is paid once per SM instead of once per tile.
// B tile : BN rows x 64 k bf16, 128 B rows, hardware 128 B swizzle (K-major).
//
// The A (activation) operand comes from im2col-mode TMA: one 128 pixel x 64 channel tile per (tap, channel
// block).  That handles every geometry (stride 2, dgrad through PixelShuffle, parity-split dgrad).  The 64 -> 64
// channel stride-1 layers take the halo-fed kernel of igemm_pm.cu instead (one TMA box per tile, pixels on M).
// Measured and removed (notes: profiles/r1_notes.md, profiles/r2_notes.md): a halo-box feed for these
// 128-pixel tiles (no gain: the thin layers sat on the per-instruction floor) and a CTA-pair kernel
// (cta_group::2, M = 256: validated, 4-10 % slower than this kernel on the 256 / 512-channel layers).
#include <stdio.h>
#include <stdlib.h>

#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace sisr {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 192;
constexpr int kATileBytes = kBM * kBK * 2;
constexpr int kMaxCout = 512;        // per-CTA bias / BN statistics arrays

struct EpiParams {
  __nv_bfloat16* out;
  int OH, OW, ldc, osy, osx, opy, opx, ps_c;
  const float* bias;
  int act;
  float slope;
  const float* slope_ptr;
  float* stats;
  int stats_rows;
  const __nv_bfloat16* mask;
  float mask_slope;
};

struct ClassTable {
  int n;                 // number of tap classes (>= 1)
  int tap_begin[4], tap_count[4], opy[4], opx[4];
};

struct KParams {
  int M, GH, GW, trav_stride, lower_w, lower_h;
  int cin_blocks, num_taps, n_tiles, m_tiles;
  ClassTable cls;
  IgemmTaps taps;
  EpiParams e;
};

// Transposed tiles (Cout <= 128): the WEIGHTS are the M operand (128 rows, rows >= Cout zero-filled by
// TMA) and 256 PIXELS are the N operand.  Measured: a tcgen05.mma with M = 128 costs >= 128 cycles per
// K = 16 step whatever N is (the A operand is fetched from shared memory row by row), so a 128 pixel x
// 64 channel instruction runs at 25 % of the tensor peak (ncu: V 64->64 @96, 514 cycles per 64-wide
// k-block, profiles/r1_notes.md).  With the pixels on the N side every instruction covers 256 pixels.
struct TParams {
  int M, GH, GW, trav_stride, lower_w, lower_h;
  int cin_blocks, num_taps, p_tiles, cout;
  ClassTable cls;
  int flat;   // output pixel index == accumulator column index (no parity / stride remap)
  IgemmTaps taps;
  EpiParams e;
};

template <int BN>
struct SmemLayout {
  static constexpr int kBTileBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
};

thread_local char g_err[256] = "";

// Butterfly "transpose-reduce": on entry lane l holds v[c] for row l, column c (32 x 32);
// on exit v[0] of lane l holds the sum over the 32 rows of column l.  31 shuffles.
__device__ __forceinline__ float column_sums_32x32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// Epilogue of one 128 x BN tile for one warp (32 accumulator rows): this thread owns output pixel
// (n_img, gh, gw) (ignored when !valid).
template <int BN>
__device__ __forceinline__ void epilogue_tile(const EpiParams& e, uint32_t tmem_tile, int quad, int lane,
                                              int n0, int cout, int n_img, int gh, int gw, bool valid,
                                              float slope, const float* s_bias, float* s_stats, int opy,
                                              int opx) {
  // element offset of this thread's pixel row for column chunk c
  auto chunk_off = [&](int c) -> size_t {
    const int ncol = n0 + c * 32;
    int oy, ox, ch;
    if (e.ps_c > 0) {
      const int sub = ncol / e.ps_c;
      ch = ncol - sub * e.ps_c;
      oy = gh * 2 + (sub >> 1);
      ox = gw * 2 + (sub & 1);
    } else {
      ch = ncol;
      oy = gh * e.osy + opy;
      ox = gw * e.osx + opx;
    }
    return (static_cast<size_t>(n_img * e.OH + oy) * e.OW + ox) * e.ldc + ch;
  };
  // fused activation backward of the tensor this gradient belongs to: its 64 mask bytes per chunk are
  // requested one chunk ahead, so their latency hides behind the TMEM read and the math of this chunk
  const bool masked = e.mask && valid;
  uint4 mnext[4];
  if (masked) {
    const uint4* m4 = reinterpret_cast<const uint4*>(e.mask + chunk_off(0));
#pragma unroll
    for (int q = 0; q < 4; ++q) mnext[q] = __ldg(m4 + q);
  }
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t raw[32];
    tmem_ld_32x32(tmem_tile + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, raw);
    uint4 mcur[4];
    if (masked) {
#pragma unroll
      for (int q = 0; q < 4; ++q) mcur[q] = mnext[q];
      if (c + 1 < BN / 32) {
        const uint4* m4 = reinterpret_cast<const uint4*>(e.mask + chunk_off(c + 1));
#pragma unroll
        for (int q = 0; q < 4; ++q) mnext[q] = __ldg(m4 + q);
      }
    }
    tmem_ld_wait();
    const int ncol = n0 + c * 32;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = __uint_as_float(raw[i]) + s_bias[ncol + i];
      if (e.act != ACT_NONE) x = x > 0.f ? x : x * slope;
      v[i] = x;
    }
    const size_t off = chunk_off(c);
    if (masked) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mcur[q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 mf = __bfloat1622float2(mh[j]);
          if (!(mf.x > 0.f)) v[q * 8 + 2 * j] *= e.mask_slope;
          if (!(mf.y > 0.f)) v[q * 8 + 2 * j + 1] *= e.mask_slope;
        }
      }
    }
    uint32_t packed[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    if (valid) {
      __nv_bfloat16* dst = e.out + off;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        d4[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
    }
    if (e.stats) {
      // statistics of the values as stored (bf16-rounded), invalid rows contribute zero
      float q[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&packed[i]);
        const float a = valid ? __low2float(h) : 0.f;
        const float b = valid ? __high2float(h) : 0.f;
        v[2 * i] = a;
        v[2 * i + 1] = b;
        q[2 * i] = a * a;
        q[2 * i + 1] = b * b;
      }
      const float cs = column_sums_32x32(v, lane);
      const float cq = column_sums_32x32(q, lane);
      atomicAdd(&s_stats[ncol + lane], cs);
      atomicAdd(&s_stats[cout + ncol + lane], cq);
    }
  }
}

// per-CTA partial BN sums -> row blockIdx.x of the stats buffer (plain stores; bn_finalize adds rows)
__device__ __forceinline__ void flush_stats(const EpiParams& e, const float* s_stats, int cout) {
  asm volatile("bar.sync 1, 128;" ::: "memory");
  float* mine = e.stats + static_cast<size_t>(blockIdx.x) * 2 * cout;
  for (int i = threadIdx.x - 64; i < 2 * cout; i += 128) mine[i] = s_stats[i];
  for (int r = gridDim.x + blockIdx.x; r < e.stats_rows; r += gridDim.x) {
    float* z = e.stats + static_cast<size_t>(r) * 2 * cout;
    for (int i = threadIdx.x - 64; i < 2 * cout; i += 128) z[i] = 0.f;
  }
}

__device__ __forceinline__ float resolve_slope(const EpiParams& e) {
  if (e.act == ACT_PRELU) return __ldg(e.slope_ptr);
  if (e.act == ACT_RELU) return 0.f;
  return e.slope;
}

// ------------------------------------------------------------------ im2col-fed kernel
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const KParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment is required by the 128 B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_stats[2 * kMaxCout];
  __shared__ float s_bias[kMaxCout];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int per_class = p.m_tiles * p.n_tiles;
  const int num_tiles = per_class * p.cls.n;
  const int cout = p.n_tiles * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 4);   // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), 2 * BN);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < cout; i += 128) s_bias[i] = p.e.bias ? p.e.bias[i] : 0.f;
    if (p.e.stats)
      for (int i = threadIdx.x - 64; i < 2 * cout; i += 128) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int hw = p.GH * p.GW;
      uint32_t kbg = 0;   // k-block counter across tiles (pipeline stage / phase)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int cls = tile / per_class, t_in = tile - cls * per_class;
        const int m0 = (t_in / p.n_tiles) * kBM;
        const int n0 = (t_in % p.n_tiles) * BN;
        const int n_img = m0 / hw;
        const int rem = m0 - n_img * hw;
        const int gh = rem / p.GW;
        const int gw = rem - gh * p.GW;
        const int cw = gw * p.trav_stride + p.lower_w;
        const int ch = gh * p.trav_stride + p.lower_h;
        for (int tap = p.cls.tap_begin[cls]; tap < p.cls.tap_begin[cls] + p.cls.tap_count[cls]; ++tap) {
          for (int cb = 0; cb < p.cin_blocks; ++cb, ++kbg) {
            const uint32_t s = kbg % STAGES;
            const uint32_t round = kbg / STAGES;
            mbar_wait(smem_u32(&empty_bar[s]), (round & 1) ^ 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, L::kStageBytes);
            const uint32_t a_dst = smem_u32(smem + s * L::kStageBytes);
            const uint32_t b_dst = a_dst + kATileBytes;
            tma_load_im2col_4d(a_dst, &tmap_a, fb, cb * kBK, cw, ch, n_img, p.taps.off_w[tap],
                               p.taps.off_h[tap]);
            tma_load_2d(b_dst, &tmap_b, fb, p.taps.k_off[tap] + cb * kBK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
    uint32_t kbg = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, use = it >> 1;
      const int num_kb = p.cls.tap_count[tile / per_class] * p.cin_blocks;
      mbar_wait(smem_u32(&tmem_empty_bar[acc]), (use & 1) ^ 1);   // epilogue drained this buffer
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
        const uint32_t s = kbg % STAGES;
        const uint32_t round = kbg / STAGES;
        mbar_wait(smem_u32(&full_bar[s]), round & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));
          if (kb == num_kb - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int hw = p.GH * p.GW;
    const float slope = resolve_slope(p.e);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, use = it >> 1;
      const int cls = tile / per_class, t_in = tile - cls * per_class;
      const int n0 = (t_in % p.n_tiles) * BN;
      const int row = (t_in / p.n_tiles) * kBM + quad * 32 + lane;
      const bool valid = row < p.M;
      const int rr = valid ? row : 0;
      const int n_img = rr / hw;
      const int rem = rr - n_img * hw;
      const int gh = rem / p.GW;
      const int gw = rem - gh * p.GW;
      mbar_wait(smem_u32(&tmem_full_bar[acc]), use & 1);
      tc_fence_after();
      epilogue_tile<BN>(p.e, tmem_base + acc * BN, quad, lane, n0, cout, n_img, gh, gw, valid, slope,
                        s_bias, s_stats, p.cls.opy[cls], p.cls.opx[cls]);
      // this warp's TMEM reads of the buffer are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
    if (p.e.stats) flush_stats(p.e, s_stats, cout);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ transposed kernel (Cout <= 128)
constexpr int kTP = 256;                          // pixels per tile (UMMA N)
constexpr int kTWBytes = 128 * kBK * 2;           // weight tile: 128 rows x 64 k
constexpr int kTPBytes = kTP * kBK * 2;           // pixel tile: 256 rows x 64 k
constexpr int kTStage = kTWBytes + kTPBytes;      // 48 KB
constexpr int kTThreads = 320;                    // TMA warp + MMA warp + 8 epilogue warps

template <int STAGES>
__global__ void __launch_bounds__(kTThreads, 1)
igemm_t_kernel(const __grid_constant__ CUtensorMap tmap_px, const __grid_constant__ CUtensorMap tmap_w,
               const TParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) uint32_t s_stage[2][32 * (64 + 16)];   // [half][32 px][cout/2 + 16 words]
  __shared__ float s_sum[2 * 128];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.p_tiles * p.cls.n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_px);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 8);   // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), 2 * kTP);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int hw = p.GH * p.GW;
      uint32_t kbg = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int cls = tile / p.p_tiles;
        const int m0 = (tile - cls * p.p_tiles) * kTP;
        const int n_img = m0 / hw;
        const int rem = m0 - n_img * hw;
        const int gh = rem / p.GW;
        const int gw = rem - gh * p.GW;
        const int cw = gw * p.trav_stride + p.lower_w;
        const int ch = gh * p.trav_stride + p.lower_h;
        for (int tap = p.cls.tap_begin[cls]; tap < p.cls.tap_begin[cls] + p.cls.tap_count[cls]; ++tap) {
          for (int cb = 0; cb < p.cin_blocks; ++cb, ++kbg) {
            const uint32_t s = kbg % STAGES;
            mbar_wait(smem_u32(&empty_bar[s]), ((kbg / STAGES) & 1) ^ 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, kTStage);
            const uint32_t w_dst = smem_u32(smem + s * kTStage);
            tma_load_2d(w_dst, &tmap_w, fb, p.taps.k_off[tap] + cb * kBK, 0);
            tma_load_im2col_4d(w_dst + kTWBytes, &tmap_px, fb, cb * kBK, cw, ch, n_img,
                               p.taps.off_w[tap], p.taps.off_h[tap]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: D[co, px] += W * X^T
    constexpr uint32_t idesc = umma_idesc_bf16(128, kTP, 0, 0);
    uint32_t kbg = 0, it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, use = it >> 1;
      const int num_kb = p.cls.tap_count[tile / p.p_tiles] * p.cin_blocks;
      mbar_wait(smem_u32(&tmem_empty_bar[acc]), (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kTP;
      for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
        const uint32_t s = kbg % STAGES;
        mbar_wait(smem_u32(&full_bar[s]), (kbg / STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t w_addr = smem_u32(smem + s * kTStage);
          const uint32_t x_addr = w_addr + kTWBytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_smem_desc(w_addr + k * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(x_addr + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));
          if (kb == num_kb - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: lane = output channel
    // TMEM row = channel, columns = pixels.  Eight warps: warp (quad, half) owns TMEM lanes
    // 32*quad.. and the 32-pixel chunks c with c % 2 == half.  Per chunk: bias / activation / bf16
    // rounding and BN sums on the thread's own channel, then a transpose through shared memory (lane
    // pairs swap one value per pixel pair so that every thread writes a packed {co, co+1} word) and
    // 16-byte NHWC stores.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int co = quad * 32 + lane;
    const EpiParams& e = p.e;
    const float slope = resolve_slope(e);
    const bool warp_active = quad * 32 < p.cout;
    const float bias = (warp_active && e.bias) ? e.bias[co] : 0.f;
    const int hw = p.GH * p.GW;
    const int active = p.cout < 128 ? 64 : 128;          // threads per half that own real channels
    const int tid_a = quad * 32 + lane;
    const int row_words = p.cout / 2 + 16;                 // staging row pitch in 32-bit words
    const int segs = p.cout / 8;                           // 16-byte segments per pixel
    if (e.stats)
      for (int i = threadIdx.x - 64; i < 2 * 128; i += 256) s_sum[i] = 0.f;
    // output pixel (flattened N*OH*OW index) of GEMM row px of tile class cls
    auto out_pixel = [&](int px, int cls) -> size_t {
      if (p.flat) return static_cast<size_t>(px);
      const int n_img = px / hw;
      const int rem = px - n_img * hw;
      const int gh = rem / p.GW;
      const int gw = rem - gh * p.GW;
      return static_cast<size_t>(n_img * e.OH + gh * e.osy + p.cls.opy[cls]) * e.OW + gw * e.osx + p.cls.opx[cls];
    };
    // fused activation backward (mask = the tensor this gradient belongs to): the four 16-byte mask words
    // a thread needs for its stores of a 32-pixel chunk are requested one chunk ahead
    uint4 mq[4];
    auto prefetch_mask = [&](int px0, int cls) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = tid_a + k * active;
        const int row = idx / segs, seg = idx - row * segs;
        const int px = px0 + row;
        if (px < p.M) mq[k] = __ldg(reinterpret_cast<const uint4*>(e.mask + out_pixel(px, cls) * e.ldc + seg * 8));
      }
    };
    float s1 = 0.f, s2 = 0.f;   // BN statistics of this channel over this warp's chunks
    uint32_t it = 0, chunk_no = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, use = it >> 1;
      const int cls = tile / p.p_tiles;
      const int m0 = (tile - cls * p.p_tiles) * kTP;
      if (e.mask && warp_active) prefetch_mask(m0 + half * 32, cls);   // lands while the MMAs of this tile run
      mbar_wait(smem_u32(&tmem_full_bar[acc]), use & 1);
      tc_fence_after();
      if (warp_active) {
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kTP;
        uint32_t raw[32];
        tmem_ld_32x32(trow + half * 32, raw);
#pragma unroll 1
        for (int c = half; c < kTP / 32; c += 2, ++chunk_no) {
          uint32_t* stage = s_stage[half];
          tmem_ld_wait();
          const int px0 = m0 + c * 32;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(raw[i]) + bias;
            if (e.act != ACT_NONE) x = x > 0.f ? x : x * slope;
            x = bf16_round(x);
            v[i] = x;
            if (px0 + i < p.M) {
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
          }
          if (c + 2 < kTP / 32) tmem_ld_32x32(trow + (c + 2) * 32, raw);   // prefetch the next chunk
          const bool odd = lane & 1;
          // the staging tile is free again once every thread finished the previous chunk's stores
          asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(active) : "memory");
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[i] : v[i + 1], 1);
            const uint32_t word = odd ? pack_bf16x2(recv, v[i + 1]) : pack_bf16x2(v[i], recv);
            stage[(i + (odd ? 1 : 0)) * row_words + (co >> 1)] = word;
          }
          asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(active) : "memory");
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int idx = tid_a + k * active;
            const int row = idx / segs, seg = idx - row * segs;
            const int px = px0 + row;
            if (px < p.M) {
              const size_t opix = out_pixel(px, cls);
              uint4 val = *reinterpret_cast<const uint4*>(stage + row * row_words + seg * 4);
              if (e.mask) {
                const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mq[k]);
                __nv_bfloat162* vh = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 mf = __bfloat1622float2(mh[j]);
                  float2 vf = __bfloat1622float2(vh[j]);
                  if (!(mf.x > 0.f)) vf.x *= e.mask_slope;
                  if (!(mf.y > 0.f)) vf.y *= e.mask_slope;
                  vh[j] = __floats2bfloat162_rn(vf.x, vf.y);
                }
              }
              *reinterpret_cast<uint4*>(e.out + opix * e.ldc + seg * 8) = val;
            }
          }
          if (e.mask && c + 2 < kTP / 32) prefetch_mask(m0 + (c + 2) * 32, cls);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
    if (e.stats) {
      asm volatile("bar.sync 3, 256;" ::: "memory");      // s_sum zeroed by all epilogue threads
      if (warp_active) {
        atomicAdd(&s_sum[co], s1);
        atomicAdd(&s_sum[128 + co], s2);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      float* mine = e.stats + static_cast<size_t>(blockIdx.x) * 2 * p.cout;
      for (int i = threadIdx.x - 64; i < 2 * p.cout; i += 256)
        mine[i] = s_sum[(i < p.cout) ? i : 128 + (i - p.cout)];
      for (int r = gridDim.x + blockIdx.x; r < e.stats_rows; r += gridDim.x) {
        float* z = e.stats + static_cast<size_t>(r) * 2 * p.cout;
        for (int i = threadIdx.x - 64; i < 2 * p.cout; i += 256) z[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kTP);
  }
}


template <typename K, typename P>
int launch_kernel(K kernel, bool* configured, int smem_bytes, const CUtensorMap& ta, const CUtensorMap& tb,
                  const P& kp, int grid, cudaStream_t stream, int threads = kThreads) {
  if (!*configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 3;
    }
    *configured = true;
  }
  kernel<<<grid, threads, smem_bytes, stream>>>(ta, tb, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "igemm launch: %s", cudaGetErrorString(e));
    return 4;
  }
  return 0;
}

template <int BN, int STAGES>
int launch_im2col(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, int grid,
                  cudaStream_t stream) {
  static bool configured = false;
  return launch_kernel(igemm_tc_kernel<BN, STAGES>, &configured, STAGES * SmemLayout<BN>::kStageBytes + 1024,
                       ta, tb, kp, grid, stream);
}
int g_transposed = 1;  // 1: Cout <= 128 layers run with the pixels on the UMMA N side
int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        g_num_sms <= 0)
      g_num_sms = 148;
  }
  return g_num_sms;
}

// Cost model (cycles): a tile costs max(operand bytes / L2->SM rate, MMA cycles) + a fixed bubble;
// the persistent grid runs ceil(tiles / SMs) of them back to back.
constexpr double kL2BytesPerClk = 48.0;
constexpr double kTileBubble = 400.0;
double plan_cost(long long tiles, double bytes_per_tile, double mma_clk) {
  const long long waves = (tiles + num_sms() - 1) / num_sms();
  const double t = bytes_per_tile / kL2BytesPerClk;
  return static_cast<double>(waves) * ((t > mma_clk ? t : mma_clk) + kTileBubble);
}

struct Plan {
  int bn;
  double cost;
};

Plan make_plan(const IgemmProblem& p) {
  Plan best{};
  best.cost = -1.0;
  const long long M = static_cast<long long>(p.NB) * p.GH * p.GW;
  const int m_tiles_i = static_cast<int>((M + kBM - 1) / kBM);
  const int cin_blocks = p.Cin / kBK;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (p.Cout % bn) continue;
    if (p.ps_c > 0 && (p.ps_c % 32)) continue;
    const int n_tiles = p.Cout / bn;
    const double mma = 2.0 * bn * 9 * cin_blocks;
    const double bytes = 9.0 * cin_blocks * (kATileBytes + bn * 128.0);
    const double c = plan_cost(static_cast<long long>(m_tiles_i) * n_tiles, bytes, mma);
    if (best.cost < 0 || c < best.cost) best = Plan{bn, c};
  }
  return best;
}

void fill_classes(ClassTable& c, const IgemmProblem& p) {
  if (p.n_classes > 1) {
    c.n = p.n_classes;
    for (int i = 0; i < 4; ++i) {
      c.tap_begin[i] = p.cls_tap_begin[i]; c.tap_count[i] = p.cls_tap_count[i];
      c.opy[i] = p.cls_opy[i]; c.opx[i] = p.cls_opx[i];
    }
  } else {
    c.n = 1;
    for (int i = 0; i < 4; ++i) {
      c.tap_begin[i] = 0; c.tap_count[i] = p.num_taps; c.opy[i] = p.opy; c.opx[i] = p.opx;
    }
  }
}

void fill_epi(EpiParams& e, const IgemmProblem& p) {
  e.out = p.out;
  e.OH = p.OH; e.OW = p.OW; e.ldc = p.ldc;
  e.osy = p.osy; e.osx = p.osx; e.opy = p.opy; e.opx = p.opx;
  e.ps_c = p.ps_c;
  e.bias = p.bias; e.act = p.act; e.slope = p.slope; e.slope_ptr = p.slope_ptr;
  e.stats = p.stats; e.stats_rows = p.stats_rows;
  e.mask = p.mask; e.mask_slope = p.mask_slope;
}

}  // namespace

const char* igemm_last_error() { return g_err; }
int igemm_max_ctas() { return num_sms(); }
void igemm_set_transposed(int on) { g_transposed = on; }

bool igemm_supported(const IgemmProblem& p) {
  if (p.Cin % 64 || p.Cout % 64 || p.Cout > kMaxCout) return false;
  if (p.stats && p.stats_rows < num_sms()) return false;
  if (p.num_taps < 1 || p.num_taps > kMaxTaps) return false;
  if (p.n_classes > 4 || (p.n_classes > 1 && p.stats)) return false;
  if (p.ldc % 8) return false;
  if (p.ps_c > 0 && (p.ps_c % 32)) return false;
  if ((reinterpret_cast<uintptr_t>(p.x) & 15) || (reinterpret_cast<uintptr_t>(p.w) & 15) ||
      (reinterpret_cast<uintptr_t>(p.out) & 15) || (reinterpret_cast<uintptr_t>(p.mask) & 15))
    return false;
  if (p.mask && p.ps_c > 0) return false;
  return true;
}

int igemm_launch(const IgemmProblem& p, cudaStream_t stream) {
  if (!igemm_supported(p)) {
    snprintf(g_err, sizeof g_err, "igemm: unsupported problem Cin=%d Cout=%d taps=%d ldc=%d", p.Cin,
             p.Cout, p.num_taps, p.ldc);
    return 1;
  }
  if (igemm_pm_supported(p)) {
    const int rc = igemm_pm_launch(p, stream);
    if (rc) snprintf(g_err, sizeof g_err, "%s", igemm_pm_last_error());
    return rc;
  }
  const bool flat_out = p.osy == 1 && p.osx == 1 && p.opy == 0 && p.opx == 0 && p.OH == p.GH && p.OW == p.GW;
  if (g_transposed && p.Cout <= 128 && p.ps_c == 0 && flat_out && p.n_classes <= 1) {
    CUtensorMap tpx, tw;
    if (make_tmap_2d_bf16(&tw, p.w, p.Cout, p.Ktot, p.Ktot, kBK, 128)) {
      snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
      return 2;
    }
    if (make_tmap_im2col_nhwc_bf16(&tpx, p.x, p.NB, p.H, p.W, p.Cin, p.lower_w, p.lower_h, p.upper_w,
                                   p.upper_h, kBK, kTP, p.trav_stride)) {
      snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
      return 2;
    }
    const long long Mt = static_cast<long long>(p.NB) * p.GH * p.GW;
    TParams tp;
    tp.M = static_cast<int>(Mt);
    tp.GH = p.GH; tp.GW = p.GW; tp.trav_stride = p.trav_stride;
    tp.lower_w = p.lower_w; tp.lower_h = p.lower_h;
    tp.cin_blocks = p.Cin / kBK;
    tp.num_taps = p.num_taps;
    tp.p_tiles = static_cast<int>((Mt + kTP - 1) / kTP);
    tp.cout = p.Cout;
    tp.taps = p.taps;
    fill_classes(tp.cls, p);
    fill_epi(tp.e, p);
    tp.flat = (p.n_classes <= 1 && p.osy == 1 && p.osx == 1 && p.opy == 0 && p.opx == 0 && p.OH == p.GH &&
               p.OW == p.GW) ? 1 : 0;
    const int t_tiles = tp.p_tiles * tp.cls.n;
    const int grid = t_tiles < num_sms() ? t_tiles : num_sms();
    static bool configured = false;
    return launch_kernel(igemm_t_kernel<4>, &configured, 4 * kTStage + 1024, tpx, tw, tp, grid, stream,
                         kTThreads);
  }
  const Plan pl = make_plan(p);
  if (pl.cost < 0) {
    snprintf(g_err, sizeof g_err, "igemm: no tile for Cout=%d", p.Cout);
    return 1;
  }
  const int bn = pl.bn;
  CUtensorMap ta, tb;
  if (make_tmap_2d_bf16(&tb, p.w, p.Cout, p.Ktot, p.Ktot, kBK, bn)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  if (make_tmap_im2col_nhwc_bf16(&ta, p.x, p.NB, p.H, p.W, p.Cin, p.lower_w, p.lower_h, p.upper_w,
                                 p.upper_h, kBK, kBM, p.trav_stride)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  const long long M = static_cast<long long>(p.NB) * p.GH * p.GW;
  KParams kp;
  kp.M = static_cast<int>(M);
  kp.GH = p.GH;
  kp.GW = p.GW;
  kp.trav_stride = p.trav_stride;
  kp.lower_w = p.lower_w;
  kp.lower_h = p.lower_h;
  kp.cin_blocks = p.Cin / kBK;
  kp.num_taps = p.num_taps;
  kp.n_tiles = p.Cout / bn;
  kp.m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  kp.taps = p.taps;
  fill_classes(kp.cls, p);
  fill_epi(kp.e, p);
  const int tiles = kp.m_tiles * kp.n_tiles * kp.cls.n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  switch (bn) {
    case 64:
      return launch_im2col<64, 8>(ta, tb, kp, grid, stream);
    case 128:
      return launch_im2col<128, 6>(ta, tb, kp, grid, stream);
    default:
      return launch_im2col<256, 4>(ta, tb, kp, grid, stream);
  }
}

}  // namespace sisr
