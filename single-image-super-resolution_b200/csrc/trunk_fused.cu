// Persistent, fused forward pass of the generator's 64-channel residual trunk (model_generator.py:5-19, 36-41,
// 86-93): ALL conv3x3 -> train-mode BatchNorm -> PReLU / skip layers of the trunk in ONE cooperative launch.
//
// Why: unfused, every layer is three launches (conv 11.8 us + statistics finalize 3.5 us + normalise 3.7 us at
// batch 64) on a 4.7 MB tensor that never leaves L2; the conv's tensor pipe is busy a quarter of its own
// duration and the rest is launch, prologue, pipeline fill and drain.  Here one CTA per SM owns the same
// R-row tile of one image in every layer (12 rows x 24 columns at 24 x 24), keeps TMEM, barriers and the
// tensor maps alive across layers, and per layer does
//   conv (the igemm_th scheme: stacked filter taps on M, pixels on N from one TMA box with halo)
//   -> epilogue A: + bias, bf16 rounding, per-channel sums; the conv output y goes to global memory (the
//      backward pass needs it) AND stays in shared memory
//   -> grid barrier #1 (the per-CTA partial sums are visible)
//   -> every CTA adds the partial rows in the same order -> batch mean / variance -> scale, shift
//      (CTA 0 also stores them for the backward pass and updates the running statistics)
//   -> pass B from shared memory: a = PReLU(scale * y + shift) or scale * y + shift + residual -> global
//   -> grid barrier #2 (the neighbours' halo rows of `a` are visible) -> next layer.
// The next layer's 72 KB of weights are fetched during pass B.  Train-mode statistics are a true batch-wide
// dependency, so two grid barriers per layer are the floor of any fused design.
// Single GPU only: with SyncBN the cross-GPU exchange would have to live inside barrier #1; the data-parallel
// path keeps the per-layer kernels.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"
#include "trunk_fused.h"

namespace sisr {

namespace {

constexpr int kThreads = 320;
constexpr int kWTile = 128 * 64 * 2;       // one stacked weight tile: 128 rows x 64 k
constexpr int kGroups = 6;                 // 3 tap pairs + 3 single taps
constexpr int kXRow = 33;                  // pitch (words) of the D_hi hand-over buffer
constexpr int kXBytes = 2 * 64 * kXRow * 4;
constexpr int kYPitch = 36;                // words per position of the y tile in shared memory (64 bf16 + pad)

struct TrunkParams {
  int NB, H, W, R, PW, tiles_h, num_tiles, n_layers;
  int n_total, chunks, chunk_n, n_valid, box_bytes, box_alloc;
  float count, momentum, eps;
  int w_row_stride;                // rows of the weight matrix between consecutive layers
  long long layer_elems;           // elements of one layer's activation tensor
  __nv_bfloat16* y_all;            // [n_layers][NB, H, W, 64]: conv outputs (kept for the backward pass)
  __nv_bfloat16* a_all;            // [n_layers][NB, H, W, 64]: layer outputs
  float* partials;                 // [n_layers][grid][128]
  unsigned int* barrier;           // zeroed before the launch
  long long* timing;               // debug (SISR_TRUNK_TIMING=1): [n_layers][8] clock64 stamps of CTA 0, else null
  TrunkLayerDev layer[kTrunkMaxLayers];
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// All CTAs are co-resident (cooperative launch).  Barrier number k (1, 2, ...): every CTA adds 1 to `count`; the
// CTA whose addition completes k * grid arrivals publishes k in `flag` (a different 128-byte line), which the
// others poll with a short sleep - the pollers never touch the line the atomics go to (with all CTAs spinning
// on the arrival counter itself a barrier cost ~5 us: 128 pollers and 128 atomics on one L2 line).
__device__ __forceinline__ void grid_barrier(unsigned int* count, unsigned int k) {
  fence_proxy_async_all();           // this thread's global stores -> later TMA (async proxy) reads elsewhere
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* flag = count + 32;
    __threadfence();
    const unsigned int old = atomicAdd(count, 1u);
    if (old + 1 == k * gridDim.x) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(k) : "memory");
    } else {
      const long long t0 = clock64();
      unsigned int seen;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (seen >= k) break;
        __nanosleep(20);
        if (clock64() - t0 > 4000000000LL) {
          printf("sisr: trunk grid barrier timeout (block %d, barrier %u, flag %u)\n", blockIdx.x, k, seen);
          __trap();
        }
      }
    }
    __threadfence();
    fence_proxy_async_all();
  }
  __syncthreads();
}

constexpr int kMaxItems = 8;      // 16-byte items of pass B per thread: n_valid * 8 <= kMaxItems * kThreads
constexpr int kMaxChunks = 5;     // 32-position chunks per epilogue half: n_valid <= 320

__global__ void __launch_bounds__(kThreads, 1)
trunk_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x0,
                 const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ TrunkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_box = smem + kGroups * kWTile;
  float* xbuf = reinterpret_cast<float*>(smem_box + p.box_alloc);          // [half][parity][64][kXRow]
  float* s_red10 = xbuf;                                                   // [10][128], statistics phase only
  uint32_t* ytile = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(xbuf) + 2 * kXBytes);   // [y_rows][kYPitch]
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t box_full;
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_sum[2 * 64];
  __shared__ float s_scale[64], s_shift[64];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_img = blockIdx.x / p.tiles_h;
  const int h0 = (blockIdx.x - n_img * p.tiles_h) * p.R;

  // zero rows: upper halves of the single-tap weight tiles (groups 3..5) and the box rows behind the TMA box
  for (int g = 3; g < kGroups; ++g)
    for (int i = threadIdx.x; i < kWTile / 2 / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(smem_w + g * kWTile + kWTile / 2)[i] = make_uint4(0, 0, 0, 0);
  for (int i = p.box_bytes / 16 + threadIdx.x; i < p.box_alloc / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_box)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x0);
    tma_prefetch_desc(&tmap_a);
    mbar_init(smem_u32(&w_bar), 1);
    mbar_init(smem_u32(&box_full), 1);
    mbar_init(smem_u32(&tmem_full_bar[0]), 1);
    mbar_init(smem_u32(&tmem_full_bar[1]), 1);
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

  // weights of layer l: rows [l * w_row_stride, + 64) of the prepared-weight matrix, tap t at columns 64 t
  auto load_weights = [&](int l) {
    const uint32_t wb = smem_u32(&w_bar);
    mbar_expect_tx(wb, 9 * (kWTile / 2));
    for (int dy = 0; dy < 3; ++dy) {
      tma_load_2d(smem_u32(smem_w + dy * kWTile), &tmap_w, wb, (dy * 3 + 0) * 64, l * p.w_row_stride);
      tma_load_2d(smem_u32(smem_w + dy * kWTile + kWTile / 2), &tmap_w, wb, (dy * 3 + 1) * 64,
                  l * p.w_row_stride);
      tma_load_2d(smem_u32(smem_w + (3 + dy) * kWTile), &tmap_w, wb, (dy * 3 + 2) * 64,
                  l * p.w_row_stride);
    }
  };
  if (warp == 0 && lane == 0) load_weights(0);

  // ---- layer-invariant geometry: the tile, its positions and this thread's share of them are the same in
  // every layer, so the index arithmetic (divisions by the box pitch) is done once
  // pass B items: item = threadIdx.x + k * kThreads -> (position o, 16-byte segment); pix < 0: not a pixel
  int b_pix[kMaxItems], b_src[kMaxItems];
#pragma unroll
  for (int k = 0; k < kMaxItems; ++k) {
    const int item = threadIdx.x + k * kThreads;
    const int o = item >> 3, seg = item & 7;
    const int r = o / p.PW, c = o - r * p.PW;
    const bool ok = item < p.n_valid * 8 && c < p.W && h0 + r < p.H;
    b_pix[k] = ok ? ((n_img * p.H + h0 + r) * p.W + c) * 8 + seg : -1;      // in 16-byte units of a [.., 64] tensor
    b_src[k] = o * kYPitch + seg * 4;
  }
  // epilogue A: validity of the 32 positions of each chunk of this warp's half
  const int quad = warp & 3;
  const int half = warp >= 2 ? (warp - 2) >> 2 : 0;
  const bool is_lo = quad < 2;
  const int co = (quad & 1) * 32 + lane;
  const int n_pchunks = (p.n_valid + 31) >> 5;
  uint32_t vm[kMaxChunks];
#pragma unroll
  for (int j = 0; j < kMaxChunks; ++j) {
    const int o = (half + 2 * j) * 32 + lane;
    const int r = o / p.PW, c = o - r * p.PW;
    vm[j] = __ballot_sync(0xffffffffu, o < p.n_valid && c < p.W && h0 + r < p.H);
  }
  float bias_next = (warp >= 2 && p.layer[0].bias) ? p.layer[0].bias[co] : 0.f;

  unsigned int barriers_done = 0;
  for (int l = 0; l < p.n_layers; ++l) {
    const TrunkLayerDev& L = p.layer[l];
    const uint32_t parity = l & 1;
    __nv_bfloat16* const y_out = p.y_all + static_cast<size_t>(l) * p.layer_elems;
    __nv_bfloat16* const a_out = p.a_all + static_cast<size_t>(l) * p.layer_elems;
    // per-channel BN parameters and the activation slope: requested now, used after barrier 1
    float bn_g = 0.f, bn_b = 0.f, bn_rm = 0.f, bn_rv = 0.f;
    if (threadIdx.x < 64) {
      bn_g = L.gamma[threadIdx.x];
      bn_b = L.beta[threadIdx.x];
      if (blockIdx.x == 0) {
        bn_rm = L.running_mean[threadIdx.x];
        bn_rv = L.running_var[threadIdx.x];
      }
    }
    const float slope = L.slope ? __ldg(L.slope) : 1.f;
    const bool stamp2 = p.timing && blockIdx.x == 0 && threadIdx.x == 64;
    // ------------------------------------------------------------------ conv
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t fb = smem_u32(&box_full);
        mbar_expect_tx(fb, p.box_bytes);
        // a_all is one [n_layers * NB, H, W, 64] tensor: layer l - 1, image n_img = index (l - 1) * NB + n_img
        if (l == 0)
          tma_load_4d(smem_u32(smem_box), &tmap_x0, fb, 0, -1, h0 - 1, n_img);
        else
          tma_load_4d(smem_u32(smem_box), &tmap_a, fb, 0, -1, h0 - 1, (l - 1) * p.NB + n_img);
      }
    } else if (warp == 1) {
      const uint32_t idesc = umma_idesc_bf16(128, p.chunk_n, 0, 0);
      mbar_wait(smem_u32(&w_bar), parity);
      mbar_wait(smem_u32(&box_full), parity);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t w_addr = smem_u32(smem_w), x_addr = smem_u32(smem_box);
#pragma unroll 1
        for (int c = 0; c < p.chunks; ++c) {
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
              const int sigma = g < 3 ? g * p.PW : (g - 3) * p.PW + 2;
              const uint64_t da = umma_smem_desc(w_addr + g * kWTile + k * 32, 16, 1024);
              const uint64_t db = umma_smem_desc(x_addr + (sigma + c * p.chunk_n) * 128 + k * 32, 16, 1024);
              umma_bf16(tmem_base + c * p.chunk_n, da, db, idesc, (k > 0 || g > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&tmem_full_bar[c]));
        }
      }
      __syncwarp();
    } else {
      // -------------------------------------------------------------- epilogue A
      // quads 2,3 (D_hi) hand their values to quads 0,1 (D_lo) through a double-buffered shared-memory tile, so
      // that they run one chunk ahead; bias, bf16 rounding, channel sums and the y tile are the lo warps' work
      const float bias = bias_next;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      float s1 = 0.f, s2 = 0.f;
      int ready = 0;
      const bool stamp = p.timing && blockIdx.x == 0 && threadIdx.x == 64;
      if (stamp) p.timing[l * 8 + 0] = clock64();
      if (threadIdx.x - 64 < 128) s_sum[threadIdx.x - 64] = 0.f;
      int j = 0;
#pragma unroll 1
      for (int pc = half; pc < n_pchunks; pc += 2, ++j) {
        const int o0 = pc * 32;
        const int par = j & 1;
        float* xb = xbuf + ((half * 2 + par) * 64) * kXRow;
        const int bar_full = 4 + half * 2 + par, bar_free = 8 + half * 2 + par;
        int need = (o0 + 32) / p.chunk_n + 1;
        if (need > p.chunks) need = p.chunks;
        while (ready < need) {
          mbar_wait(smem_u32(&tmem_full_bar[ready]), parity);
          ++ready;
        }
        tc_fence_after();
        if (stamp && j == 0) p.timing[l * 8 + 1] = clock64();       // first instruction chunk complete
        uint32_t raw[32];
        tmem_ld_32x32(trow + o0 + (is_lo ? 0 : 1), raw);
        tmem_ld_wait();
        if (!is_lo) {
          if (j >= 2) asm volatile("bar.sync %0, 128;" ::"r"(bar_free) : "memory");   // chunk j-2 has been read
#pragma unroll
          for (int i = 0; i < 32; ++i) xb[co * kXRow + i] = __uint_as_float(raw[i]);
          asm volatile("bar.arrive %0, 128;" ::"r"(bar_full) : "memory");
        } else {
          const uint32_t vmask = vm[j < kMaxChunks ? j : kMaxChunks - 1];
          asm volatile("bar.sync %0, 128;" ::"r"(bar_full) : "memory");
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = bf16_round(__uint_as_float(raw[i]) + xb[co * kXRow + i] + bias);
            v[i] = x;
            if ((vmask >> i) & 1u) {
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
          }
          asm volatile("bar.arrive %0, 128;" ::"r"(bar_free) : "memory");
          // transposed write into the y tile: a lane pair swaps one value per position pair, every thread
          // writes packed {co, co+1} words (the tile keeps the whole layer output of this CTA)
          const bool odd = lane & 1;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[i] : v[i + 1], 1);
            const uint32_t word = odd ? pack_bf16x2(recv, v[i + 1]) : pack_bf16x2(v[i], recv);
            ytile[(o0 + i + (odd ? 1 : 0)) * kYPitch + (co >> 1)] = word;
          }
        }
      }
      if (!is_lo) {
        // consume the "read" arrivals of the last two chunks, so that every named barrier ends the layer balanced
        if (j >= 2) asm volatile("bar.sync %0, 128;" ::"r"(8 + half * 2 + (j & 1)) : "memory");
        if (j >= 1) asm volatile("bar.sync %0, 128;" ::"r"(8 + half * 2 + ((j - 1) & 1)) : "memory");
      }
      while (ready < p.chunks) {
        mbar_wait(smem_u32(&tmem_full_bar[ready]), parity);
        ++ready;
      }
      tc_fence_before();
      if (is_lo) {
        atomicAdd(&s_sum[co], s1);
        atomicAdd(&s_sum[64 + co], s2);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");       // y tile and the CTA's sums are complete
      if (stamp) p.timing[l * 8 + 2] = clock64();
      const int et = threadIdx.x - 64;                      // 0..255
      if (et < 128) p.partials[(static_cast<size_t>(l) * gridDim.x + blockIdx.x) * 128 + et] = s_sum[et];
      if (l + 1 < p.n_layers) bias_next = p.layer[l + 1].bias ? p.layer[l + 1].bias[co] : 0.f;
    }
    if (stamp2) p.timing[l * 8 + 3] = clock64();
    grid_barrier(p.barrier, ++barriers_done);
    if (stamp2) p.timing[l * 8 + 4] = clock64();

    // ------------------------------------------------------------------ statistics -> scale / shift
    // residual rows of pass B: requested now, consumed after the statistics
    const uint4* res4 = reinterpret_cast<const uint4*>(L.residual);
    uint4 rvk[kMaxItems];
#pragma unroll
    for (int k = 0; k < kMaxItems; ++k) {
      rvk[k] = make_uint4(0, 0, 0, 0);
      if (res4 && b_pix[k] >= 0) rvk[k] = res4[b_pix[k]];
    }
    {
      // partial rows [grid][128]: warp w adds rows w, w + 10, ... (one coalesced 512-byte row per warp load, up to
      // 16 independent loads in flight per thread), then the ten warp sums are added in a fixed order - the
      // same bits on every CTA
      const float4* src = reinterpret_cast<const float4*>(p.partials + static_cast<size_t>(l) * gridDim.x * 128) + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const int nrows = static_cast<int>(gridDim.x);
#pragma unroll 1
      for (int r0 = warp; r0 < nrows; r0 += 10 * 16) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int r = r0 + u * 10;
          v[u] = r < nrows ? __ldcg(src + static_cast<size_t>(r) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
        }
      }
      reinterpret_cast<float4*>(s_red10 + warp * 128)[lane] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int c = threadIdx.x;
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int w10 = 0; w10 < 10; ++w10) {
        sum += s_red10[w10 * 128 + c];
        sq += s_red10[w10 * 128 + 64 + c];
      }
      const float mean = sum / p.count;
      const float var = fmaxf(sq / p.count - mean * mean, 0.f);
      const float invstd = rsqrtf(var + p.eps);
      const float sc = bn_g * invstd;
      const float sh = bn_b - mean * sc;
      s_scale[c] = sc;
      s_shift[c] = sh;
      if (blockIdx.x == 0) {
        L.aux[c] = sc;
        L.aux[64 + c] = sh;
        L.aux[128 + c] = mean;
        L.aux[192 + c] = invstd;
        const float unbiased = p.count > 1.f ? var * p.count / (p.count - 1.f) : var;
        L.running_mean[c] = (1.f - p.momentum) * bn_rm + p.momentum * mean;
        L.running_var[c] = (1.f - p.momentum) * bn_rv + p.momentum * unbiased;
        if (c == 0 && L.nbt) *L.nbt += 1;
      }
    }
    __syncthreads();
    if (warp == 0 && lane == 0 && l + 1 < p.n_layers) load_weights(l + 1);    // lands during pass B / barrier 2

    if (stamp2) p.timing[l * 8 + 5] = clock64();
    // ------------------------------------------------------------------ pass B: y -> global; normalise (+PReLU / +residual)
    {
      uint4* const y4 = reinterpret_cast<uint4*>(y_out);
      uint4* const a4 = reinterpret_cast<uint4*>(a_out);
#pragma unroll
      for (int k = 0; k < kMaxItems; ++k) {
        if (b_pix[k] < 0) continue;
        uint4 val = *reinterpret_cast<const uint4*>(ytile + b_src[k]);
        y4[b_pix[k]] = val;                              // the conv output, kept for the backward pass
        const int seg = (threadIdx.x + k * kThreads) & 7;
        __nv_bfloat162* vh = reinterpret_cast<__nv_bfloat162*>(&val);
        const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rvk[k]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int ch = seg * 8 + 2 * jj;
          float2 f = __bfloat1622float2(vh[jj]);
          f.x = fmaf(f.x, s_scale[ch], s_shift[ch]);
          f.y = fmaf(f.y, s_scale[ch + 1], s_shift[ch + 1]);
          if (L.slope) {
            f.x = f.x > 0.f ? f.x : f.x * slope;
            f.y = f.y > 0.f ? f.y : f.y * slope;
          }
          if (res4) {
            const float2 rr = __bfloat1622float2(rh[jj]);
            f.x += rr.x;
            f.y += rr.y;
          }
          vh[jj] = __floats2bfloat162_rn(f.x, f.y);
        }
        a4[b_pix[k]] = val;
      }
    }
    if (stamp2) p.timing[l * 8 + 6] = clock64();
    grid_barrier(p.barrier, ++barriers_done);
    if (stamp2) p.timing[l * 8 + 7] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_sms = 0;
int sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}
thread_local char g_err[256] = "";

struct Plan {
  int R, n_total, chunks, chunk_n, box_alloc, smem, tiles_h, tiles;
};

bool make_plan(int nb, int h, int w, Plan& best) {
  const int PW = w + 2;
  double best_cost = -1.0;
  if (PW > 256) return false;
  for (int R = 1; R <= h && R + 2 <= 256; ++R) {
    int n_total = (R * PW + 15) / 16 * 16, chunks = 1;
    if (n_total > 256) {
      n_total = (R * PW + 31) / 32 * 32;
      chunks = 2;
    }
    if (n_total > 512 || R * PW > 32 * 2 * kMaxChunks || R * PW * 8 > kMaxItems * kThreads) break;
    const int chunk_n = n_total / chunks;
    const int rows_needed = n_total + 2 * PW + 2 > (R + 2) * PW ? n_total + 2 * PW + 2 : (R + 2) * PW;
    const int box_alloc = (rows_needed * 128 + 1023) / 1024 * 1024;
    const int tiles_h = (h + R - 1) / R;
    const long long tiles = static_cast<long long>(nb) * tiles_h;
    if (tiles > sms()) continue;                        // every tile needs its own co-resident CTA
    const int y_rows = (R * PW + 31) / 32 * 32;          // the epilogue writes whole 32-position chunks
    const int smem = kGroups * kWTile + box_alloc + 2 * kXBytes + y_rows * kYPitch * 4 + 1024;
    if (smem > 225 * 1024) continue;
    const double mma = 4.0 * kGroups * chunks * (chunk_n * 0.5 + 38.0 > 94.0 ? chunk_n * 0.5 + 38.0 : 94.0);
    const double cost = (mma + 14.0 * n_total + 1500.0) * (1.0 + 0.02 * (tiles_h * R - h));
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = Plan{R, n_total, chunks, chunk_n, box_alloc, smem, tiles_h, static_cast<int>(tiles)};
    }
  }
  return best_cost >= 0;
}

}  // namespace

const char* trunk_fused_last_error() { return g_err; }

bool trunk_fused_supported(int nb, int h, int w, int n_layers) {
  Plan pl;
  int coop = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  return coop && n_layers >= 1 && n_layers <= kTrunkMaxLayers && make_plan(nb, h, w, pl);
}

size_t trunk_fused_workspace_bytes(int n_layers) {
  return sizeof(float) * static_cast<size_t>(n_layers) * sms() * 128 + 256;
}

int trunk_fused_forward(const __nv_bfloat16* x0, int nb, int h, int w, const __nv_bfloat16* weights,
                        int w_row_stride, const TrunkLayerHost* layers, int n_layers, __nv_bfloat16* y_all,
                        __nv_bfloat16* a_all, float momentum, float eps, void* workspace, cudaStream_t stream) {
  Plan pl;
  if (!trunk_fused_supported(nb, h, w, n_layers) || !make_plan(nb, h, w, pl)) {
    snprintf(g_err, sizeof g_err, "trunk_fused: unsupported geometry %d x %d x %d, %d layers", nb, h, w, n_layers);
    return 1;
  }
  const int PW = w + 2;
  CUtensorMap tw, tx0, ta;
  if (make_tmap_2d_bf16(&tw, weights, static_cast<uint64_t>(n_layers - 1) * w_row_stride + 64, 576, 576, 64, 64) ||
      make_tmap_tiled_nhwc_bf16(&tx0, x0, nb, h, w, 64, 64, PW, pl.R + 2) ||
      make_tmap_tiled_nhwc_bf16(&ta, a_all, n_layers * nb, h, w, 64, 64, PW, pl.R + 2)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  TrunkParams p{};
  p.NB = nb; p.H = h; p.W = w; p.R = pl.R; p.PW = PW;
  p.tiles_h = pl.tiles_h; p.num_tiles = pl.tiles; p.n_layers = n_layers;
  p.n_total = pl.n_total; p.chunks = pl.chunks; p.chunk_n = pl.chunk_n;
  p.n_valid = pl.R * PW;
  p.box_bytes = (pl.R + 2) * PW * 128;
  p.box_alloc = pl.box_alloc;
  p.count = static_cast<float>(nb) * h * w;
  p.momentum = momentum; p.eps = eps;
  p.w_row_stride = w_row_stride;
  p.layer_elems = static_cast<long long>(nb) * h * w * 64;
  p.y_all = y_all; p.a_all = a_all;
  p.barrier = static_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  for (int l = 0; l < n_layers; ++l) {
    const TrunkLayerHost& s = layers[l];
    TrunkLayerDev& d = p.layer[l];
    d.bias = s.bias; d.gamma = s.gamma; d.beta = s.beta;
    d.running_mean = s.running_mean; d.running_var = s.running_var; d.nbt = s.nbt;
    d.slope = s.slope; d.aux = s.aux;
    if (s.residual_layer >= l) {
      snprintf(g_err, sizeof g_err, "trunk_fused: layer %d takes its residual from a later layer", l);
      return 1;
    }
    d.residual = s.residual_layer == -2 ? x0
               : (s.residual_layer >= 0 ? a_all + static_cast<size_t>(s.residual_layer) * p.layer_elems : nullptr);
  }
  static int configured = 0;
  if (pl.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(trunk_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 3;
    }
    configured = pl.smem;
  }
  cudaMemsetAsync(workspace, 0, 256, stream);
  static const bool want_timing = getenv("SISR_TRUNK_TIMING") != nullptr;      // debug: phase breakdown of CTA 0
  static long long* d_timing = nullptr;
  if (want_timing && !d_timing) cudaMalloc(&d_timing, sizeof(long long) * kTrunkMaxLayers * 8);
  p.timing = want_timing ? d_timing : nullptr;
  void* args[] = {&tw, &tx0, &ta, &p};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(trunk_fwd_kernel), dim3(pl.tiles),
                                              dim3(kThreads), args, pl.smem, stream);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "trunk_fused launch: %s", cudaGetErrorString(e));
    return 4;
  }
  if (want_timing) {
    static int printed = 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    if (cap == cudaStreamCaptureStatusNone && printed < 2 && cudaStreamSynchronize(stream) == cudaSuccess) {
      ++printed;
      long long h[kTrunkMaxLayers * 8];
      cudaMemcpy(h, d_timing, sizeof(long long) * n_layers * 8, cudaMemcpyDeviceToHost);
      double acc[8] = {};
      for (int l = 1; l < n_layers; ++l) {
        acc[0] += h[l * 8 + 0] - h[(l - 1) * 8 + 7];      // after barrier 2 -> epilogue warps enter
        for (int i = 1; i < 8; ++i) acc[i] += h[l * 8 + i] - h[l * 8 + i - 1];
      }
      const char* names[8] = {"layer entry", "box load + first MMA chunk", "epilogue A (+ second chunk)",
                              "partials + y store", "grid barrier 1", "statistics", "pass B", "grid barrier 2"};
      double tot = 0;
      for (int i = 0; i < 8; ++i) tot += acc[i];
      printf("trunk_fused phase breakdown (CTA 0, cycles per layer, %d layers):\n", n_layers - 1);
      for (int i = 0; i < 8; ++i) printf("  %-30s %8.0f\n", names[i], acc[i] / (n_layers - 1));
      printf("  %-30s %8.0f\n", "total", tot / (n_layers - 1));
      fflush(stdout);
    }
  }
  return 0;
}

}  // namespace sisr
