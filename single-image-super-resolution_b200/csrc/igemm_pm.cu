// Halo-fed tcgen05 convolution for the 64 -> 64 channel 3x3 layers (stride 1, same size) with the PIXELS on
// the UMMA M side: the generator trunk (fprop and dgrad: 66 launches per step) and VGG conv1_2.
//
// Why not weights on M with two stacked taps and the pixels on N (round 2's first scheme, igemm_th, removed)?
// Its MMA stream is short (24 full instructions per 320 positions) but its epilogue is not: the accumulator
// holds channels on the TMEM lanes, so the upper lane half (the stacked tap) must be handed to the lower half
// through shared memory, every value crosses a lane-pair transpose and a staging tile before it can leave as
// an NHWC row, and the BN sums are taken per lane.  With the pixels on M
//   * an accumulator row IS an NHWC pixel: each epilogue thread reads 32 channels of its own position from
//     TMEM and stores 64 contiguous bytes - no hand-over, no transpose, no staging tile, no named barriers;
//   * one M tile = 128 consecutive box positions x 64 output channels = 64 TMEM columns, so up to 8 tiles
//     (1024 positions) fit the 512 columns: the epilogue of one tile runs under the MMAs of the others;
//   * the nine taps are nine start-address shifts of the A descriptor into ONE TMA box with halo
//     (csrc/probe_shift.cu); the weights (72 KB, nine 64 x 64 K-major tiles) stay resident in shared memory
//     for all tiles of the CTA.
// Measured with the phase counters below (profiles/r2_pm_phases.txt): the kernel is bound by the MMA stream -
// the issuing thread is held ~107 cycles per M = 128 x N = 64 x K = 16 instruction (36 per M tile), the epilogue
// warps work 300-500 cycles per M tile and wait for the rest.  A CTA-pair form (cta_group::2, M = 256, each CTA
// its own tile and half of the weight rows) was built and passed parity, but its instruction took 186 cycles in
// this kernel - slower than two single-CTA instructions - and was removed (profiles/r2_notes.md).
// The two halo columns of every box row are computed and discarded (<= 8 %).
//
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-9: epilogue (warp % 4 = TMEM lane quadrant = 32
//   positions of the M tile, (warp - 2) / 4 = channel half)
#include <stdio.h>
#include <stdlib.h>

#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace sisr {

namespace {

constexpr int kThreads = 320;
constexpr int kWTap = 64 * 64 * 2;         // one tap: 64 output channels (rows) x 64 input channels, 128 B rows
constexpr int kWBytes = 9 * kWTap;
constexpr int kMaxMT = 8;                  // M tiles per box (64 TMEM columns each)

struct PMParams {
  int H, W, R, TW, PW, tiles_h, tiles_w, num_tiles;   // tile = R rows x TW columns, box pitch PW = TW + 2
  int mt;                           // M tiles (128 box positions each) per tile
  int sets;                         // accumulator sets (2: the epilogue of tile i runs under the MMAs of tile i+1)
  int grp;                          // M tiles whose instructions are interleaved
  int box_bytes, box_alloc, nbox;
  int sigma[9];                     // box-row shift of the pixel operand per tap
  int k_off[9];                     // first weight column of the tap
  // epilogue
  __nv_bfloat16* out;
  int ldc;
  const float* bias;
  int act;
  float slope;
  const float* slope_ptr;
  float* stats;
  int stats_rows;
  const __nv_bfloat16* mask;   // nullable, indexed like `out`: stored value = mask > 0 ? v : v * mask_slope
  float mask_slope;
  long long* dbg;              // nullable (harness only): phase cycle counters of CTA 0, see igemm_set_pm_debug
};

thread_local char g_err[256] = "";
long long* g_pm_dbg = nullptr;

// v[0..32) per lane -> lane l holds the sum over the warp of element l (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(kThreads, 1)
igemm_pm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const PMParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_box = smem + kWBytes;
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t box_full[2];
  __shared__ __align__(8) uint64_t box_empty[2];
  __shared__ __align__(8) uint64_t tmem_full_bar[2][kMaxMT];    // [accumulator set][M tile]
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_bias[64];
  __shared__ float s_part[8][64];            // per epilogue warp: {sum[32], sum of squares[32]} of its channel half

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long k0 = p.dbg ? clock64() : 0;

  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    mbar_init(smem_u32(&w_bar), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&box_full[i]), 1);
      mbar_init(smem_u32(&box_empty[i]), 1);
      for (int m = 0; m < kMaxMT; ++m) mbar_init(smem_u32(&tmem_full_bar[i][m]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 8);   // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t wb = smem_u32(&w_bar);
      mbar_expect_tx(wb, kWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(smem_w + t * kWTap), &tmap_w, wb, p.k_off[t], 0);
      // Programmatic dependent launch: barriers, TMEM and the weights (prepared long before the preceding
      // kernel) overlap the tail of the kernel that produces the input tensor; the activation boxes wait for
      // its completion.  Every global write of this kernel happens after an MMA that consumed such a box.
      pdl_wait();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t bb = it % p.nbox, use = it / p.nbox;
        mbar_wait(smem_u32(&box_empty[bb]), (use & 1) ^ 1);
        const int per_img = p.tiles_h * p.tiles_w;
        const int n_img = tile / per_img;
        const int t2 = tile - n_img * per_img;
        const int h0 = (t2 / p.tiles_w) * p.R, w0 = (t2 % p.tiles_w) * p.TW;
        const uint32_t fb = smem_u32(&box_full[bb]);
        mbar_expect_tx(fb, p.box_bytes);
        tma_load_4d(smem_u32(smem_box + bb * p.box_alloc), &tmap_x, fb, 0, w0 - 1, h0 - 1, n_img);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: D[position, co] += X_tap * W_tap^T
    {
      // (descriptor arithmetic and the 108 instructions of a tile are issued by one lane)
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      mbar_wait(smem_u32(&w_bar), 0);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t set = it % p.sets, suse = it / p.sets;
        const uint32_t bb = it % p.nbox, buse = it / p.nbox;
        const long long c0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&tmem_empty_bar[set]), (suse & 1) ^ 1);   // epilogue(s) drained this accumulator set
        const long long c1 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&box_full[bb]), buse & 1);
        const long long c2 = p.dbg ? clock64() : 0;
        tc_fence_after();
        if (lane == 0) {
          const uint32_t tmem_d = tmem_base + set * p.mt * 64;
          const uint32_t w_addr = smem_u32(smem_w);
          const uint32_t x_addr = smem_u32(smem_box + bb * p.box_alloc);
          auto issue = [&](int m, int k, int t) {
            const uint64_t db = umma_smem_desc(w_addr + t * kWTap + k * 32, 16, 1024);
            const uint64_t da = umma_smem_desc(x_addr + (m * 128 + p.sigma[t]) * 128 + k * 32, 16, 1024);
            umma_bf16(tmem_d + m * 64, da, db, idesc, (k > 0 || t > 0) ? 1u : 0u);
          };
          auto commit = [&](uint32_t bar) { umma_commit(bar); };
#pragma unroll 1
          for (int m0 = 0; m0 < p.mt; m0 += p.grp) {
            const int m1 = m0 + p.grp < p.mt ? m0 + p.grp : p.mt;
            // K steps 0-2: consecutive instructions go to different M tiles of the group; last K step: tile by
            // tile, so that the first tiles of the group are complete (and in the epilogue) while the
            // instructions of the later ones still run
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
#pragma unroll 1
                for (int m = m0; m < m1; ++m) issue(m, k, t);
              }
            }
#pragma unroll 1
            for (int m = m0; m < m1; ++m) {
#pragma unroll
              for (int t = 0; t < 9; ++t) issue(m, 3, t);
              commit(smem_u32(&tmem_full_bar[set][m]));
            }
          }
          commit(smem_u32(&box_empty[bb]));
          if (p.dbg && blockIdx.x == 0) {
            p.dbg[0] += c1 - c0;             // MMA warp waits for a free accumulator set
            p.dbg[1] += c2 - c1;             // ... for the box
            p.dbg[2] += clock64() - c2;      // ... issues
            p.dbg[3] += 1;
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    float slope = p.slope;
    if (p.act == ACT_PRELU) slope = __ldg(p.slope_ptr);
    if (p.act == ACT_RELU) slope = 0.f;
    float s1[32], s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s1[i] = s2[i] = 0.f;
    uint32_t it = 0;
    long long t_prev = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t set = it % p.sets, suse = it / p.sets;
      const int per_img = p.tiles_h * p.tiles_w;
      const int n_img = tile / per_img;
      const int t2 = tile - n_img * per_img;
      const int h0 = (t2 / p.tiles_w) * p.R, w0 = (t2 % p.tiles_w) * p.TW;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + set * p.mt * 64 + half * 32;
#pragma unroll 1
      for (int m = 0; m < p.mt; ++m) {
        // this lane's position of the M tile: output pixel and validity
        const int o = m * 128 + quad * 32 + lane;
        const int r = o / p.PW, c = o - r * p.PW;
        const bool ok = r < p.R && c < p.TW && w0 + c < p.W && h0 + r < p.H;
        const size_t opix = ok ? static_cast<size_t>((n_img * p.H + h0 + r) * p.W + w0 + c) : 0;
        __nv_bfloat16* dst = p.out + opix * p.ldc + half * 32;
        uint4 mk[4];
        if (p.mask && ok) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask + opix * p.ldc + half * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) mk[j] = __ldg(mp + j);
        }
        const long long e0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&tmem_full_bar[set][m]), suse & 1);
        if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) {
          const long long e1 = clock64();
          p.dbg[4] += e1 - e0;               // epilogue warp 2 waits for an M tile
          if (t_prev) p.dbg[5] += e0 - t_prev;   // ... works on the previous one (incl. its stores being issued)
          p.dbg[6] += 1;
        }
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(trow + m * 64, raw);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0 = __uint_as_float(raw[i]) + s_bias[half * 32 + i];
          float x1 = __uint_as_float(raw[i + 1]) + s_bias[half * 32 + i + 1];
          if (p.act != ACT_NONE) {
            x0 = x0 > 0.f ? x0 : x0 * slope;
            x1 = x1 > 0.f ? x1 : x1 * slope;
          }
          x0 = bf16_round(x0);
          x1 = bf16_round(x1);
          if (ok) {
            s1[i] += x0;
            s2[i] = fmaf(x0, x0, s2[i]);
            s1[i + 1] += x1;
            s2[i + 1] = fmaf(x1, x1, s2[i + 1]);
          }
          packed[i >> 1] = pack_bf16x2(x0, x1);
        }
        if (ok) {
          if (p.mask) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mk[j]);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 mf = __bfloat1622float2(mh[q]);
                float2 vf = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&packed[j * 4 + q]));
                if (!(mf.x > 0.f)) vf.x *= p.mask_slope;
                if (!(mf.y > 0.f)) vf.y *= p.mask_slope;
                packed[j * 4 + q] = pack_bf16x2(vf.x, vf.y);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            reinterpret_cast<uint4*>(dst)[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        if (p.dbg) t_prev = clock64();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[set]));
    }
    if (p.stats) {
      // per warp: lane l <- sum over the warp's 32 positions of channel l of its half; then the four
      // quadrant warps of a half are added in a fixed order (no atomics: bit-reproducible)
      const float a = warp_transpose_sum(s1, lane);
      const float b = warp_transpose_sum(s2, lane);
      s_part[warp - 2][lane] = a;
      s_part[warp - 2][32 + lane] = b;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int e = threadIdx.x - 64;                  // 0 .. 255
      if (e < 128) {
        const int which = e >> 6, ch = e & 63;         // 0: sums, 1: sums of squares
        const int hf = ch >> 5, cl = ch & 31;
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) t += s_part[hf * 4 + q][which * 32 + cl];
        p.stats[static_cast<size_t>(blockIdx.x) * 128 + e] = t;
      }
      for (int rr = gridDim.x + blockIdx.x; rr < p.stats_rows; rr += gridDim.x) {
        float* z = p.stats + static_cast<size_t>(rr) * 128;
        for (int i = e; i < 128; i += 256) z[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[7] += clock64() - k0;
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_pm_mode = [] { const char* e = getenv("SISR_PM"); return e && e[0] == '0' ? 0 : 1; }();
int g_pm_grp = [] { const char* e = getenv("SISR_PM_GRP"); return e ? atoi(e) : 0; }();     // 0: planner's choice
int g_sms = 0;
int sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

struct PMPlan {
  int R, TW, mt, sets, grp, box_alloc, nbox, smem;
  double cost;
};

// Tile = R rows x TW columns = mt M tiles of 128 box positions: minimise rounds x (MMA cycles + fixed) over the
// shapes that fit shared memory (resident weights + one or two boxes) and TMEM.
bool make_plan(const IgemmProblem& p, PMPlan& best) {
  best.cost = -1.0;
  const long long imgs = p.NB;
  for (int split = 1; split <= 4; ++split) {
    const int TW = (p.W + split - 1) / split;
    if (TW + 2 > 256 || (split > 1 && TW < 16)) continue;
    const int PW = TW + 2;
    const int tiles_w = (p.W + TW - 1) / TW;
    for (int R = 1; R <= p.H && R + 2 <= 256; ++R) {
      const int mt = (R * PW + 127) / 128;
      if (mt > kMaxMT) break;
      const int rows_needed = mt * 128 + 2 * PW + 2 > (R + 2) * PW ? mt * 128 + 2 * PW + 2 : (R + 2) * PW;
      const int box_alloc = (rows_needed * 128 + 1023) / 1024 * 1024;
      const int tiles_h = (p.H + R - 1) / R;
      const long long tiles = imgs * tiles_h * tiles_w;
      const long long rounds = (tiles + sms() - 1) / sms();        // tiles per CTA
      const int nbox = rounds > 1 ? 2 : 1;
      const int smem = kWBytes + nbox * box_alloc + 1024;
      if (smem > 224 * 1024) continue;               // + 2.5 KB of static shared memory <= 227 KB per CTA
      const int sets = (rounds > 1 && 2 * mt <= kMaxMT) ? 2 : 1;
      // 36 instructions per M tile, ~107 cycles each (measured).  With one accumulator set the epilogue of the
      // last M tile of a tile is exposed.
      const double per_instr = 107.0;
      const double epi = sets == 2 ? 0.0 : 700.0;
      const double c = static_cast<double>(rounds) * (36.0 * mt * per_instr + epi + 2500.0);
      if (best.cost < 0 || c < best.cost) best = PMPlan{R, TW, mt, sets, mt < 4 ? mt : 4, box_alloc, nbox, smem, c};
    }
  }
  return best.cost >= 0;
}

}  // namespace

const char* igemm_pm_last_error() { return g_err; }
void igemm_set_pm(int on) { g_pm_mode = on; }
void igemm_set_pm_grp(int grp) { g_pm_grp = grp; }
void igemm_set_pm_debug(long long* counters) { g_pm_dbg = counters; }

bool igemm_pm_supported(const IgemmProblem& p) {
  if (!g_pm_mode) return false;
  if (p.Cin != 64 || p.Cout != 64 || p.num_taps != 9 || p.n_classes > 1 || p.ps_c != 0) return false;
  if (p.trav_stride != 1 || p.GH != p.H || p.GW != p.W || p.lower_w != -1 || p.lower_h != -1) return false;
  if (p.osy != 1 || p.osx != 1 || p.opy != 0 || p.opx != 0 || p.OH != p.GH || p.OW != p.GW) return false;
  if (p.ldc % 8) return false;
  if (p.stats && p.stats_rows < sms()) return false;
  if (static_cast<long long>(p.NB) * p.H * p.W >= (1ll << 31)) return false;
  int seen = 0;
  for (int t = 0; t < 9; ++t) {
    if (p.taps.off_w[t] > 2 || p.taps.off_h[t] > 2) return false;
    seen |= 1 << (p.taps.off_h[t] * 3 + p.taps.off_w[t]);
  }
  if (seen != 0x1FF) return false;
  PMPlan pl;
  return make_plan(p, pl);
}

int igemm_pm_launch(const IgemmProblem& p, cudaStream_t stream) {
  PMPlan pl;
  if (!igemm_pm_supported(p) || !make_plan(p, pl)) {
    snprintf(g_err, sizeof g_err, "igemm_pm: unsupported problem");
    return 1;
  }
  CUtensorMap tx, tw;
  const int PW = pl.TW + 2;
  if (make_tmap_2d_bf16(&tw, p.w, p.Cout, p.Ktot, p.Ktot, 64, 64) ||
      make_tmap_tiled_nhwc_bf16(&tx, p.x, p.NB, p.H, p.W, p.Cin, 64, PW, pl.R + 2)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  PMParams tp;
  tp.H = p.H; tp.W = p.W; tp.R = pl.R; tp.TW = pl.TW; tp.PW = PW;
  tp.tiles_h = (p.H + pl.R - 1) / pl.R;
  tp.tiles_w = (p.W + pl.TW - 1) / pl.TW;
  tp.num_tiles = p.NB * tp.tiles_h * tp.tiles_w;
  tp.mt = pl.mt; tp.sets = pl.sets;
  tp.grp = g_pm_grp > 0 ? (g_pm_grp < pl.mt ? g_pm_grp : pl.mt) : pl.grp;
  tp.box_bytes = (pl.R + 2) * PW * 128;
  tp.box_alloc = pl.box_alloc; tp.nbox = pl.nbox;
  for (int t = 0; t < 9; ++t) {
    tp.sigma[t] = p.taps.off_h[t] * PW + p.taps.off_w[t];
    tp.k_off[t] = p.taps.k_off[t];
  }
  tp.out = p.out; tp.ldc = p.ldc;
  tp.bias = p.bias; tp.act = p.act; tp.slope = p.slope; tp.slope_ptr = p.slope_ptr;
  tp.stats = p.stats; tp.stats_rows = p.stats_rows;
  tp.mask = p.mask; tp.mask_slope = p.mask_slope;
  tp.dbg = g_pm_dbg;
  static int configured = 0;
  if (pl.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_pm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 3;
    }
    configured = pl.smem;
  }
  const int grid = tp.num_tiles < sms() ? tp.num_tiles : sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, igemm_pm_kernel, tx, tw, tp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "igemm_pm launch: %s", cudaGetErrorString(e));
    return 4;
  }
  return 0;
}

}  // namespace sisr
