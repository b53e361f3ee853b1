// Transposed, halo-fed tcgen05 convolution for the 64 -> 64 channel 3x3 layers (stride 1, same size):
// the generator trunk (fprop and dgrad: 66 launches per step) and VGG conv1_2.
//
// What the round-2 hardware probe says (profiles/r2_probe_mma.txt, cycles per K = 16 tcgen05.mma, both
// operands in shared memory): M=128,N=64: 94   M=128,N=128: 103   M=128,N=256: 167   (a ~94-cycle floor per
// instruction, then ~0.5 cycle per N column).  A 64-channel layer therefore wants
//   * the PIXELS on the N side (N up to 256 per instruction) and the weights on the M side,
//   * all 128 M rows doing useful work: two filter taps are STACKED along M (rows 0-63 = tap A, rows 64-127 =
//     tap B) against ONE pixel operand.  With the B operand shifted by tap A's offset, the lower rows
//     accumulate y[o] += W_A x[o + s_A] and the upper rows W_B x[o + s_A], which is tap B's contribution to
//     output position o + s_A - s_B.  Pairing (dy,0) with (dy,1) makes that difference -1 for every pair, so
//     y[o] = D_lo[o] + D_hi[o + 1]: one shifted add in the epilogue (a TMEM column offset).  The three
//     (dy,2) taps run with zero upper rows: 6 instructions per K step instead of 9 half-empty ones.
//   * ONE tiled TMA box per tile (R image rows + halo, pitch PW = W + 2, zero fill outside the image =
//     the conv padding) instead of nine im2col tiles: the nine taps are nine start-address shifts of the
//     shared-memory descriptor into that box (csrc/probe_shift.cu), 9x less L2 -> SM traffic, and the
//     six stacked weight tiles (96 KB) stay resident in shared memory for all tiles of the CTA.
// Accumulator column n = box position r * PW + c; the two halo columns of every row are computed and
// discarded (<= 8 %).  N > 256 positions are issued as two instructions per (tap group, K step).
//
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-9: epilogue (quads 0,1 own D_lo = channels 0-63,
//   quads 2,3 own D_hi and hand it over through shared memory; bias / activation / bf16 rounding / BN sums on
//   the thread's own channel, lane-pair transpose, 16-byte NHWC stores)
#include <stdio.h>
#include <stdlib.h>

#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace sisr {

namespace {

constexpr int kThreads = 320;
constexpr int kWTile = 128 * 64 * 2;       // one stacked weight tile: 128 rows x 64 k
constexpr int kGroups = 6;                 // 3 tap pairs + 3 single taps
constexpr int kXRow = 33;                  // pitch (words) of the D_hi hand-over buffer
constexpr int kXBytes = 2 * 64 * kXRow * 4;            // [half][64 channels][32 positions + pad]
constexpr int kStageWords = 32 * (32 + 16);            // staging tile of one half: 32 px x (cout/2 + 16) words
constexpr int kStageBytes = 2 * kStageWords * 4;

struct THParams {
  int H, W, R, TW, PW, tiles_h, tiles_w, num_tiles;   // tile = R rows x TW columns, box pitch PW = TW + 2
  int n_total, chunks, chunk_n;     // accumulator columns, MMA instructions per group, columns per instruction
  int n_valid;                      // R * PW: positions that can hold a real pixel
  int box_bytes, box_alloc, nbox, dbuf;
  int sigma[kGroups];               // box-row shift of the pixel operand
  int k_lo[kGroups], k_hi[kGroups]; // first weight column of the lower / upper tap (k_hi < 0: single tap)
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
  float mask_slope;            // (ReLU backward of the tensor this data gradient belongs to, fused into the store)
};

thread_local char g_err[256] = "";

__global__ void __launch_bounds__(kThreads, 1)
igemm_th_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const THParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_box = smem + kGroups * kWTile;
  float* xbuf = reinterpret_cast<float*>(smem_box + p.nbox * p.box_alloc);
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(xbuf) + kXBytes);
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t box_full[2];
  __shared__ __align__(8) uint64_t box_empty[2];
  __shared__ __align__(8) uint64_t tmem_full_bar[2][2];    // [accumulator][instruction chunk]
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_sum[2 * 64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // upper halves of the single-tap weight tiles: zero rows (never written by TMA)
  for (int g = 0; g < kGroups; ++g)
    if (p.k_hi[g] < 0)
      for (int i = threadIdx.x; i < kWTile / 2 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_w + g * kWTile + kWTile / 2)[i] = make_uint4(0, 0, 0, 0);
  // rows of the box buffers behind the TMA box: the instructions of the single taps multiply them by the
  // zero weight rows (position R*PW-2 reads one row past the box), so they must hold finite values
  for (int b = 0; b < p.nbox; ++b)
    for (int i = p.box_bytes / 16 + threadIdx.x; i < p.box_alloc / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(smem_box + b * p.box_alloc)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 2 * 64) s_sum[threadIdx.x] = 0.f;
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    mbar_init(smem_u32(&w_bar), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&box_full[i]), 1);
      mbar_init(smem_u32(&box_empty[i]), 1);
      mbar_init(smem_u32(&tmem_full_bar[i][0]), 1);
      mbar_init(smem_u32(&tmem_full_bar[i][1]), 1);
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
      int n_w = 0;
      for (int g = 0; g < kGroups; ++g) n_w += p.k_hi[g] >= 0 ? 2 : 1;
      const uint32_t wb = smem_u32(&w_bar);
      mbar_expect_tx(wb, n_w * (kWTile / 2));
      for (int g = 0; g < kGroups; ++g) {
        tma_load_2d(smem_u32(smem_w + g * kWTile), &tmap_w, wb, p.k_lo[g], 0);
        if (p.k_hi[g] >= 0) tma_load_2d(smem_u32(smem_w + g * kWTile + kWTile / 2), &tmap_w, wb, p.k_hi[g], 0);
      }
      // Programmatic dependent launch: everything above (barriers, TMEM, zero fill, the 72 KB of weights -
      // prepared long before the preceding kernel) overlaps the tail of the kernel that produces the input
      // tensor; the activation boxes wait for its completion.  Every global write of this kernel happens
      // after an MMA that consumed such a box, i.e. after this wait.  (No-op for an ordinary launch.)
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
    // ------------------------------------------------------------ MMA issuer: D[row, position] += W * X^T
    const uint32_t idesc = umma_idesc_bf16(128, p.chunk_n, 0, 0);
    mbar_wait(smem_u32(&w_bar), 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = p.dbuf ? (it & 1) : 0, ause = p.dbuf ? (it >> 1) : it;
      const uint32_t bb = it % p.nbox, buse = it / p.nbox;
      mbar_wait(smem_u32(&tmem_empty_bar[acc]), (ause & 1) ^ 1);   // epilogue drained this accumulator
      mbar_wait(smem_u32(&box_full[bb]), buse & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t tmem_d = tmem_base + acc * p.n_total;
        const uint32_t w_addr = smem_u32(smem_w);
        const uint32_t x_addr = smem_u32(smem_box + bb * p.box_alloc);
        // chunk-major: the epilogue starts on the first half of the positions while the second half is
        // still being multiplied
#pragma unroll 1
        for (int c = 0; c < p.chunks; ++c) {
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
              const uint64_t da = umma_smem_desc(w_addr + g * kWTile + k * 32, 16, 1024);
              const uint64_t db = umma_smem_desc(x_addr + (p.sigma[g] + c * p.chunk_n) * 128 + k * 32, 16, 1024);
              umma_bf16(tmem_d + c * p.chunk_n, da, db, idesc, (k > 0 || g > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&tmem_full_bar[acc][c]));
        }
        umma_commit(smem_u32(&box_empty[bb]));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                 // position chunks pc with pc % 2 == half
    const bool is_lo = quad < 2;                      // lanes 0-63: D_lo, lanes 64-127: D_hi
    const int co = (quad & 1) * 32 + lane;            // output channel of this thread
    const int tid_a = co;                             // index among the 64 storing threads of the half
    float slope = p.slope;
    if (p.act == ACT_PRELU) slope = __ldg(p.slope_ptr);
    if (p.act == ACT_RELU) slope = 0.f;
    const float bias = p.bias ? p.bias[co] : 0.f;
    float* xb = xbuf + half * 64 * kXRow;
    uint32_t* stage = s_stage + half * kStageWords;
    constexpr int row_words = 32 + 16;
    const int n_pchunks = (p.n_valid + 31) >> 5;
    const int bar_full = 4 + half, bar_free = 6 + half, bar_stage = 1 + half;
    float s1 = 0.f, s2 = 0.f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = p.dbuf ? (it & 1) : 0, ause = p.dbuf ? (it >> 1) : it;
      const int per_img = p.tiles_h * p.tiles_w;
      const int n_img = tile / per_img;
      const int t2 = tile - n_img * per_img;
      const int h0 = (t2 / p.tiles_w) * p.R, w0 = (t2 % p.tiles_w) * p.TW;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * p.n_total;
      int ready = 0;                                   // instruction chunks of this accumulator known complete
#pragma unroll 1
      for (int pc = half; pc < n_pchunks; pc += 2) {
        const int o0 = pc * 32;
        int need = (o0 + 32) / p.chunk_n + 1;         // D_hi reads one column past the 32 positions
        if (need > p.chunks) need = p.chunks;
        while (ready < need) {
          mbar_wait(smem_u32(&tmem_full_bar[acc][ready]), ause & 1);
          ++ready;
        }
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(trow + o0 + (is_lo ? 0 : 1), raw);
        // this lane's position of the chunk: output pixel and validity (same in every warp of the half)
        const int o = o0 + lane;
        const int r = o / p.PW, c = o - r * p.PW;
        const bool ok = o < p.n_valid && c < p.TW && w0 + c < p.W && h0 + r < p.H;
        const int opix = ok ? (n_img * p.H + h0 + r) * p.W + w0 + c : -1;
        const uint32_t vmask = __ballot_sync(0xffffffffu, ok);
        tmem_ld_wait();
        if (!is_lo) {
#pragma unroll
          for (int i = 0; i < 32; ++i) xb[co * kXRow + i] = __uint_as_float(raw[i]);
          asm volatile("bar.arrive %0, 128;" ::"r"(bar_full) : "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(bar_free) : "memory");      // lo threads have read the buffer
        } else {
          asm volatile("bar.sync %0, 128;" ::"r"(bar_full) : "memory");
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(raw[i]) + xb[co * kXRow + i] + bias;
            if (p.act != ACT_NONE) x = x > 0.f ? x : x * slope;
            x = bf16_round(x);
            v[i] = x;
            if ((vmask >> i) & 1u) {
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
          }
          asm volatile("bar.arrive %0, 128;" ::"r"(bar_free) : "memory");
          const bool odd = lane & 1;
          // the staging tile is free again once both storing warps finished the previous chunk's stores
          asm volatile("bar.sync %0, 64;" ::"r"(bar_stage) : "memory");
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[i] : v[i + 1], 1);
            const uint32_t word = odd ? pack_bf16x2(recv, v[i + 1]) : pack_bf16x2(v[i], recv);
            stage[(i + (odd ? 1 : 0)) * row_words + (co >> 1)] = word;
          }
          asm volatile("bar.sync %0, 64;" ::"r"(bar_stage) : "memory");
          int pixk[4];
          uint4 mk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int idx = tid_a + k * 64;
            pixk[k] = __shfl_sync(0xffffffffu, opix, idx >> 3);
            if (p.mask && pixk[k] >= 0)       // the four mask words of this thread's stores: loads issued together
              mk[k] = __ldg(reinterpret_cast<const uint4*>(p.mask + static_cast<size_t>(pixk[k]) * p.ldc + (idx & 7) * 8));
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int idx = tid_a + k * 64;
            const int row = idx >> 3, seg = idx & 7;          // 8 16-byte segments per 64-channel pixel
            if (pixk[k] >= 0) {
              uint4 val = *reinterpret_cast<const uint4*>(stage + row * row_words + seg * 4);
              if (p.mask) {
                const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mk[k]);
                __nv_bfloat162* vh = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 mf = __bfloat1622float2(mh[j]);
                  float2 vf = __bfloat1622float2(vh[j]);
                  if (!(mf.x > 0.f)) vf.x *= p.mask_slope;
                  if (!(mf.y > 0.f)) vf.y *= p.mask_slope;
                  vh[j] = __floats2bfloat162_rn(vf.x, vf.y);
                }
              }
              *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(pixk[k]) * p.ldc + seg * 8) = val;
            }
          }
        }
      }
      while (ready < p.chunks) {                       // (keeps the barrier phases in step)
        mbar_wait(smem_u32(&tmem_full_bar[acc][ready]), ause & 1);
        ++ready;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
    if (p.stats) {
      if (is_lo) {
        atomicAdd(&s_sum[co], s1);
        atomicAdd(&s_sum[64 + co], s2);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      float* mine = p.stats + static_cast<size_t>(blockIdx.x) * 128;
      for (int i = threadIdx.x - 64; i < 128; i += 256) mine[i] = s_sum[i];
      for (int rr = gridDim.x + blockIdx.x; rr < p.stats_rows; rr += gridDim.x) {
        float* z = p.stats + static_cast<size_t>(rr) * 128;
        for (int i = threadIdx.x - 64; i < 128; i += 256) z[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_th_mode = [] { const char* e = getenv("SISR_TH"); return e && e[0] == '0' ? 0 : 1; }();
int g_sms = 0;
int sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

struct THPlan {
  int R, TW, n_total, chunks, chunk_n, box_alloc, nbox, dbuf, smem;
  double cost;
};

// Tile = R rows x TW columns: minimise waves x (MMA cycles + epilogue + fixed) over the shapes that fit shared
// memory / TMEM.  Full-width tiles first (one box row = one image row); column tiles when an image row is too
// wide for two box buffers beside the resident weights (VGG conv1_2 at 96 x 96).
bool make_plan(const IgemmProblem& p, THPlan& best) {
  best.cost = -1.0;
  const long long imgs = p.NB;
  const int fixed = kGroups * kWTile + kXBytes + kStageBytes + 1024;
  for (int split = 1; split <= 4; ++split) {
    const int TW = (p.W + split - 1) / split;
    if (TW + 2 > 256 || (split > 1 && TW < 16)) continue;
    const int PW = TW + 2;
    const int tiles_w = (p.W + TW - 1) / TW;
    for (int R = 1; R <= p.H && R + 2 <= 256; ++R) {
      int n_total = (R * PW + 15) / 16 * 16;
      int chunks = 1;
      if (n_total > 256) {
        n_total = (R * PW + 31) / 32 * 32;
        chunks = 2;
      }
      if (n_total > 512) break;
      const int chunk_n = n_total / chunks;
      const int rows_needed = n_total + 2 * PW + 2 > (R + 2) * PW ? n_total + 2 * PW + 2 : (R + 2) * PW;
      const int box_alloc = (rows_needed * 128 + 1023) / 1024 * 1024;
      const int tiles_h = (p.H + R - 1) / R;
      const long long tiles = imgs * tiles_h * tiles_w;
      const long long waves = (tiles + sms() - 1) / sms();
      // several tiles per CTA need two box buffers (the load of tile i+1 under the MMAs of tile i): with one
      // buffer VGG conv1_2 (96 x 96, 21 tiles per CTA) measured 130 us against 120 us on the im2col-fed kernel
      const int nbox = waves > 1 ? 2 : 1;
      if (fixed + nbox * box_alloc > 225 * 1024) continue;
      if (waves > 1 && n_total < 192) continue;          // short instructions sit on the 94-cycle floor
      const double mma = 4.0 * kGroups * chunks * (chunk_n * 0.5 + 38.0 > 94.0 ? chunk_n * 0.5 + 38.0 : 94.0);
      const double epi = 14.0 * n_total;                 // TMEM reads, hand-over, transpose, stores
      const double waste = 1.0 + 0.02 * (tiles_h * R - p.H) + 0.02 * (tiles_w * TW - p.W);
      const double c = static_cast<double>(waves) * (mma + epi + 1500.0) * waste;
      if (best.cost < 0 || c < best.cost)
        best = THPlan{R, TW, n_total, chunks, chunk_n, box_alloc, nbox, 2 * n_total <= 512 ? 1 : 0,
                      fixed + nbox * box_alloc, c};
    }
  }
  return best.cost >= 0;
}

}  // namespace

const char* igemm_th_last_error() { return g_err; }
void igemm_set_th(int on) { g_th_mode = on; }

// same-size stride-1 3x3 conv, 64 -> 64 channels, plain NHWC output, no fused gradient mask
bool igemm_th_supported(const IgemmProblem& p) {
  if (!g_th_mode) return false;
  if (p.Cin != 64 || p.Cout != 64 || p.num_taps != 9 || p.n_classes > 1 || p.ps_c != 0) return false;
  if (p.mask && p.stats) return false;
  if (p.trav_stride != 1 || p.GH != p.H || p.GW != p.W || p.lower_w != -1 || p.lower_h != -1) return false;
  if (p.osy != 1 || p.osx != 1 || p.opy != 0 || p.opx != 0 || p.OH != p.GH || p.OW != p.GW) return false;
  if (p.ldc % 8) return false;
  if (p.stats && p.stats_rows < sms()) return false;
  int seen = 0;
  for (int t = 0; t < 9; ++t) {
    if (p.taps.off_w[t] > 2 || p.taps.off_h[t] > 2) return false;
    seen |= 1 << (p.taps.off_h[t] * 3 + p.taps.off_w[t]);
  }
  if (seen != 0x1FF) return false;
  THPlan pl;
  return make_plan(p, pl);
}

int igemm_th_launch(const IgemmProblem& p, cudaStream_t stream) {
  THPlan pl;
  if (!igemm_th_supported(p) || !make_plan(p, pl)) {
    snprintf(g_err, sizeof g_err, "igemm_th: unsupported problem");
    return 1;
  }
  CUtensorMap tx, tw;
  const int PW = pl.TW + 2;
  if (make_tmap_2d_bf16(&tw, p.w, p.Cout, p.Ktot, p.Ktot, 64, 64) ||
      make_tmap_tiled_nhwc_bf16(&tx, p.x, p.NB, p.H, p.W, p.Cin, 64, PW, pl.R + 2)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  THParams tp;
  tp.H = p.H; tp.W = p.W; tp.R = pl.R; tp.TW = pl.TW; tp.PW = PW;
  tp.tiles_h = (p.H + pl.R - 1) / pl.R;
  tp.tiles_w = (p.W + pl.TW - 1) / pl.TW;
  tp.num_tiles = p.NB * tp.tiles_h * tp.tiles_w;
  tp.n_total = pl.n_total; tp.chunks = pl.chunks; tp.chunk_n = pl.chunk_n;
  tp.n_valid = pl.R * PW;
  tp.box_bytes = (pl.R + 2) * PW * 128;
  tp.box_alloc = pl.box_alloc; tp.nbox = pl.nbox; tp.dbuf = pl.dbuf;
  // tap (dy, dx) -> index in the problem's tap table
  int at[3][3];
  for (int t = 0; t < 9; ++t) at[p.taps.off_h[t]][p.taps.off_w[t]] = t;
  for (int dy = 0; dy < 3; ++dy) {
    tp.sigma[dy] = dy * PW;                         // pair: (dy,0) below, (dy,1) above -> y[o] = lo[o] + hi[o+1]
    tp.k_lo[dy] = p.taps.k_off[at[dy][0]];
    tp.k_hi[dy] = p.taps.k_off[at[dy][1]];
    tp.sigma[3 + dy] = dy * PW + 2;                 // single: (dy,2)
    tp.k_lo[3 + dy] = p.taps.k_off[at[dy][2]];
    tp.k_hi[3 + dy] = -1;
  }
  tp.out = p.out; tp.ldc = p.ldc;
  tp.bias = p.bias; tp.act = p.act; tp.slope = p.slope; tp.slope_ptr = p.slope_ptr;
  tp.stats = p.stats; tp.stats_rows = p.stats_rows;
  tp.mask = p.mask; tp.mask_slope = p.mask_slope;
  static int configured = 0;
  if (pl.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_th_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 3;
    }
    configured = pl.smem;
  }
  const int grid = tp.num_tiles < sms() ? tp.num_tiles : sms();
  static const int pdl = [] { const char* e = getenv("SISR_TH_PDL"); return e && e[0] == '0' ? 0 : 1; }();
  if (pdl) {
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
    cudaLaunchKernelEx(&cfg, igemm_th_kernel, tx, tw, tp);
  } else {
    igemm_th_kernel<<<grid, kThreads, pl.smem, stream>>>(tx, tw, tp);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "igemm_th launch: %s", cudaGetErrorString(e));
    return 4;
  }
  return 0;
}

}  // namespace sisr
