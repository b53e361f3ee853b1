// tcgen05 implicit-GEMM convolution engine for sm_100a.
//
//   warp 0      : TMA producer  (im2col-mode loads of the activation tile, tiled loads of weights)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  : epilogue (tcgen05.ld -> bias / activation / BN statistics -> bf16 NHWC store)
//
// Persistent: one CTA per SM walks the 128 x BN output tiles (tile = blockIdx.x + i * gridDim.x);
// the K loop of a tile runs over (tap, 64-channel block).  The TMEM accumulator is double
// buffered (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i+1 and
// barrier / TMEM set-up is paid once per SM instead of once per tile.
// A tile : 128 pixels x 64 channels bf16 = 16 KB, 128 B rows, hardware 128 B swizzle (K-major).
// B tile : BN rows x 64 k bf16, same layout.  Accumulator: 128 lanes x BN fp32 columns in TMEM.
#include <stdio.h>

#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"

namespace sisr {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 192;
constexpr int kATileBytes = kBM * kBK * 2;

constexpr int kMaxCout = 512;   // per-CTA BN statistics accumulator

struct KParams {
  int M, GH, GW, trav_stride, lower_w, lower_h;
  int cin_blocks, num_taps, n_tiles, m_tiles;
  // tiled A loads (stride-1 same-size convs): an M tile is a TH x TW pixel rectangle of one image,
  // fetched per tap as ONE tiled-mode TMA box {64 ch, TW, TH} shifted by the tap offset (zero fill
  // outside the image = conv padding); tiled = 0: im2col-mode TMA over 128 consecutive positions
  int tiled, TW, TH, tiles_w, tiles_hw, a_bytes;
  IgemmTaps taps;
  __nv_bfloat16* out;
  int OH, OW, ldc, osy, osx, opy, opx, ps_c;
  const float* bias;
  int act;
  float slope;
  const float* slope_ptr;
  float* stats;
  int stats_rows;
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
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int num_kb = p.num_taps * p.cin_blocks;
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
    for (int i = threadIdx.x - 64; i < cout; i += 128) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (p.stats)
      for (int i = threadIdx.x - 64; i < 2 * cout; i += 128) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int hw = p.GH * p.GW;
      uint32_t kbg = 0;   // k-block counter across tiles (pipeline stage / phase)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tile_m = tile / p.n_tiles;
        const int n0 = (tile % p.n_tiles) * BN;
        int n_img, gh, gw;
        if (p.tiled) {
          n_img = tile_m / p.tiles_hw;
          const int r2 = tile_m - n_img * p.tiles_hw;
          gh = (r2 / p.tiles_w) * p.TH;
          gw = (r2 % p.tiles_w) * p.TW;
        } else {
          const int m0 = tile_m * kBM;
          n_img = m0 / hw;
          const int rem = m0 - n_img * hw;
          gh = rem / p.GW;
          gw = rem - gh * p.GW;
        }
        const int cw = gw * p.trav_stride + p.lower_w;
        const int ch = gh * p.trav_stride + p.lower_h;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          for (int cb = 0; cb < p.cin_blocks; ++cb, ++kbg) {
            const uint32_t s = kbg % STAGES;
            const uint32_t round = kbg / STAGES;
            mbar_wait(smem_u32(&empty_bar[s]), (round & 1) ^ 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, p.a_bytes + L::kBTileBytes);
            const uint32_t a_dst = smem_u32(smem + s * L::kStageBytes);
            const uint32_t b_dst = a_dst + kATileBytes;
            if (p.tiled)
              tma_load_4d(a_dst, &tmap_a, fb, cb * kBK, cw + p.taps.off_w[tap], ch + p.taps.off_h[tap],
                          n_img);
            else
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
    float slope = p.slope;
    if (p.act == ACT_PRELU) slope = __ldg(p.slope_ptr);
    if (p.act == ACT_RELU) slope = 0.f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, use = it >> 1;
      const int tile_m = tile / p.n_tiles;
      const int n0 = (tile % p.n_tiles) * BN;
      int n_img, gh, gw;
      bool valid;
      if (p.tiled) {
        const int r = quad * 32 + lane;
        n_img = tile_m / p.tiles_hw;
        const int r2 = tile_m - n_img * p.tiles_hw;
        const int th = r / p.TW;
        gh = (r2 / p.tiles_w) * p.TH + th;
        gw = (r2 % p.tiles_w) * p.TW + (r - th * p.TW);
        valid = th < p.TH && gh < p.GH && gw < p.GW;
        if (!valid) gh = gw = 0;
      } else {
        const int row = tile_m * kBM + quad * 32 + lane;
        valid = row < p.M;
        const int rr = valid ? row : 0;
        n_img = rr / hw;
        const int rem = rr - n_img * hw;
        gh = rem / p.GW;
        gw = rem - gh * p.GW;
      }

      mbar_wait(smem_u32(&tmem_full_bar[acc]), use & 1);
      tc_fence_after();

#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + c * 32, raw);
        tmem_ld_wait();
        const int ncol = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(raw[i]) + s_bias[ncol + i];
          if (p.act != ACT_NONE) x = x > 0.f ? x : x * slope;
          v[i] = x;
        }
        // destination of this 32-channel chunk
        int oy, ox, ch;
        if (p.ps_c > 0) {
          const int sub = ncol / p.ps_c;
          ch = ncol - sub * p.ps_c;
          oy = gh * 2 + (sub >> 1);
          ox = gw * 2 + (sub & 1);
        } else {
          ch = ncol;
          oy = gh * p.osy + p.opy;
          ox = gw * p.osx + p.opx;
        }
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        if (valid) {
          __nv_bfloat16* dst =
              p.out + (static_cast<size_t>(n_img * p.OH + oy) * p.OW + ox) * p.ldc + ch;
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            d4[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
        if (p.stats) {
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
      // this warp's TMEM reads of the buffer are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
    if (p.stats) {
      // per-CTA partial sums, plain stores (no contended atomics); bn_finalize adds the rows
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* mine = p.stats + static_cast<size_t>(blockIdx.x) * 2 * cout;
      for (int i = threadIdx.x - 64; i < 2 * cout; i += 128) mine[i] = s_stats[i];
      for (int r = gridDim.x + blockIdx.x; r < p.stats_rows; r += gridDim.x) {
        float* z = p.stats + static_cast<size_t>(r) * 2 * cout;
        for (int i = threadIdx.x - 64; i < 2 * cout; i += 128) z[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, int STAGES>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, int grid,
                   cudaStream_t stream) {
  const int smem_bytes = STAGES * SmemLayout<BN>::kStageBytes + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BN, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 3;
    }
    configured = true;
  }
  igemm_tc_kernel<BN, STAGES><<<grid, kThreads, smem_bytes, stream>>>(ta, tb, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "igemm launch: %s", cudaGetErrorString(e));
    return 4;
  }
  return 0;
}

bool g_force_im2col = true;    // tiled-mode A boxes measured no faster than im2col mode (same L2 traffic)
                               // and their tile counts quantise worse; igemm_force_im2col(0) enables them
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

// Tile width: minimise waves x (MMA cycles of one tile + per-tile bubble); ties go to the wider
// tile, which re-reads the activation operand fewer times.
int pick_bn(int m_tiles, int cout, int ps_c, int num_kb) {
  const int cands[3] = {256, 128, 64};
  long long best_cost = 0;
  int best = 0;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (cout % bn) continue;
    if (ps_c > 0 && (ps_c % 32)) continue;
    const long long tiles = static_cast<long long>(m_tiles) * (cout / bn);
    const long long waves = (tiles + num_sms() - 1) / num_sms();
    const long long cost = waves * (2LL * bn * num_kb + 300);
    if (best == 0 || cost < best_cost) {
      best = bn;
      best_cost = cost;
    }
  }
  return best;
}

}  // namespace

const char* igemm_last_error() { return g_err; }
int igemm_max_ctas() { return num_sms(); }
void igemm_force_im2col(int on) { g_force_im2col = on != 0; }

bool igemm_supported(const IgemmProblem& p) {
  if (p.Cin % 64 || p.Cout % 64 || p.Cout > kMaxCout) return false;
  if (p.stats && p.stats_rows < num_sms()) return false;
  if (p.num_taps < 1 || p.num_taps > kMaxTaps) return false;
  if (p.ldc % 8) return false;
  if (p.ps_c > 0 && (p.ps_c % 32)) return false;
  if ((reinterpret_cast<uintptr_t>(p.x) & 15) || (reinterpret_cast<uintptr_t>(p.w) & 15) ||
      (reinterpret_cast<uintptr_t>(p.out) & 15))
    return false;
  return true;
}

int igemm_launch(const IgemmProblem& p, cudaStream_t stream) {
  if (!igemm_supported(p)) {
    snprintf(g_err, sizeof g_err, "igemm: unsupported problem Cin=%d Cout=%d taps=%d ldc=%d", p.Cin,
             p.Cout, p.num_taps, p.ldc);
    return 1;
  }
  const long long M = static_cast<long long>(p.NB) * p.GH * p.GW;
  int m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  // rectangular M tiles + tiled-mode TMA when a rectangle covers the image with >= 85 % useful rows
  int tw = 0, th = 0, tiles_w = 0, tiles_h = 0;
  if (p.trav_stride == 1 && p.GH == p.H && p.GW == p.W && !g_force_im2col) {
    const int cand_w[7] = {p.W <= 128 ? p.W : 0, 128, 64, 32, 16, 8, 4};
    double best = 0.0;
    for (int i = 0; i < 7; ++i) {
      const int cw = cand_w[i];
      if (cw <= 0 || cw > p.W) continue;
      const int chh = 128 / cw;
      const int nw = (p.W + cw - 1) / cw, nh = (p.H + chh - 1) / chh;
      const double eff = static_cast<double>(p.H) * p.W / (static_cast<double>(nw) * nh * 128.0);
      if (eff > best + 1e-9) {
        best = eff; tw = cw; th = chh; tiles_w = nw; tiles_h = nh;
      }
    }
    if (best < 0.85) tw = 0;
  }
  if (tw) m_tiles = p.NB * tiles_h * tiles_w;
  const int bn = pick_bn(m_tiles, p.Cout, p.ps_c, p.num_taps * (p.Cin / kBK));
  if (!bn) {
    snprintf(g_err, sizeof g_err, "igemm: no tile for Cout=%d", p.Cout);
    return 1;
  }
  CUtensorMap ta, tb;
  if (tw ? make_tmap_tiled_nhwc_bf16(&ta, p.x, p.NB, p.H, p.W, p.Cin, kBK, tw, th)
         : make_tmap_im2col_nhwc_bf16(&ta, p.x, p.NB, p.H, p.W, p.Cin, p.lower_w, p.lower_h, p.upper_w,
                                      p.upper_h, kBK, kBM, p.trav_stride)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
  if (make_tmap_2d_bf16(&tb, p.w, p.Cout, p.Ktot, p.Ktot, kBK, bn)) {
    snprintf(g_err, sizeof g_err, "%s", tmap_last_error());
    return 2;
  }
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
  kp.m_tiles = m_tiles;
  kp.tiled = tw ? 1 : 0;
  kp.TW = tw; kp.TH = th; kp.tiles_w = tiles_w; kp.tiles_hw = tiles_w * tiles_h;
  kp.a_bytes = tw ? tw * th * kBK * 2 : kATileBytes;
  kp.taps = p.taps;
  kp.out = p.out;
  kp.OH = p.OH;
  kp.OW = p.OW;
  kp.ldc = p.ldc;
  kp.osy = p.osy;
  kp.osx = p.osx;
  kp.opy = p.opy;
  kp.opx = p.opx;
  kp.ps_c = p.ps_c;
  kp.bias = p.bias;
  kp.act = p.act;
  kp.slope = p.slope;
  kp.slope_ptr = p.slope_ptr;
  kp.stats = p.stats;
  kp.stats_rows = p.stats_rows;
  const int tiles = m_tiles * kp.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  switch (bn) {
    case 64:
      return launch_variant<64, 8>(ta, tb, kp, grid, stream);
    case 128:
      return launch_variant<128, 6>(ta, tb, kp, grid, stream);
    default:
      return launch_variant<256, 4>(ta, tb, kp, grid, stream);
  }
}

}  // namespace sisr
