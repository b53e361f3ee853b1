// Stand-alone GPU harness (no Python):
//   1. probes the im2col-mode TMA traversal / padding / stride semantics the engine relies on,
//   2. checks the tcgen05 implicit-GEMM engine against the CUDA-core convolution,
//   3. times the SRGAN trunk conv shape.
// Build: see build.py (target "harness").  Run on a B200: ./harness_igemm
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "conv_simt.h"
#include "igemm.h"
#include "ptx.cuh"
#include "tmap.h"

using namespace sisr;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// ------------------------------------------------------------------ im2col probe
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, int c, int w, int h, int n,
                             int off_w, int off_h, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 128 * 128 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0xFFFFFFFFu;
  __syncthreads();
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar), 128 * 128);
    tma_load_im2col_4d(smem_u32(smem), &tmap, smem_u32(&bar), c, w, h, n, (uint16_t)off_w,
                       (uint16_t)off_h);
  }
  mbar_wait(smem_u32(&bar), 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) out[i] = smem[i];
}

struct ProbeCase {
  const char* name;
  int N, H, W, C, lower, upper, stride, GH, GW;
  int start;  // linear start position
  int off_w, off_h;
};

static int run_probe(const ProbeCase& pc) {
  const size_t elems = (size_t)pc.N * pc.H * pc.W * pc.C;
  std::vector<uint16_t> hx(elems);
  for (int n = 0; n < pc.N; ++n)
    for (int h = 0; h < pc.H; ++h)
      for (int w = 0; w < pc.W; ++w)
        for (int c = 0; c < pc.C; ++c) {
          float v = (float)(c % 7);
          if (c == 0) v = n + 1;
          if (c == 1) v = h + 1;
          if (c == 2) v = w + 1;
          if (c == 8) v = 77;
          hx[(((size_t)n * pc.H + h) * pc.W + w) * pc.C + c] = f2bf(v);
        }
  uint16_t* dx;
  uint8_t* dout;
  CK(cudaMalloc(&dx, elems * 2));
  CK(cudaMalloc(&dout, 128 * 128));
  CK(cudaMemcpy(dx, hx.data(), elems * 2, cudaMemcpyHostToDevice));
  CUtensorMap tm;
  if (make_tmap_im2col_nhwc_bf16(&tm, dx, pc.N, pc.H, pc.W, pc.C, pc.lower, pc.lower, pc.upper,
                                 pc.upper, 64, 128, pc.stride)) {
    printf("[probe %s] tensor map error: %s\n", pc.name, tmap_last_error());
    return 1;
  }
  const int hw = pc.GH * pc.GW;
  const int n0 = pc.start / hw, rem = pc.start % hw, gh0 = rem / pc.GW, gw0 = rem % pc.GW;
  probe_kernel<<<1, 128, 128 * 128 + 1024>>>(tm, 0, gw0 * pc.stride + pc.lower,
                                              gh0 * pc.stride + pc.lower, n0, pc.off_w, pc.off_h,
                                              dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[probe %s] kernel error: %s\n", pc.name, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<uint8_t> ho(128 * 128);
  CK(cudaMemcpy(ho.data(), dout, 128 * 128, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < 128; ++r) {
    // de-swizzle: logical 16 B chunk j lives at chunk j ^ (r & 7)
    auto elem = [&](int c) {
      const int chunk = (c / 8) ^ (r & 7);
      uint16_t v;
      memcpy(&v, &ho[r * 128 + chunk * 16 + (c % 8) * 2], 2);
      return bf2f(v);
    };
    const int p = pc.start + r;
    const int n = p / hw, q = p % hw, gh = q / pc.GW, gw = q % pc.GW;
    const int ih = gh * pc.stride + pc.lower + pc.off_h;
    const int iw = gw * pc.stride + pc.lower + pc.off_w;
    const bool inb = n < pc.N && ih >= 0 && ih < pc.H && iw >= 0 && iw < pc.W;
    const float en = inb ? n + 1 : 0, eh = inb ? ih + 1 : 0, ew = inb ? iw + 1 : 0,
                e8 = inb ? 77 : 0;
    const float gn = elem(0), gh_ = elem(1), gw_ = elem(2), g8 = elem(8);
    const bool ok = gn == en && gh_ == eh && gw_ == ew && g8 == e8;
    if (!ok) ++bad;
    if (!ok && bad <= 12)
      printf("[probe %s] row %3d: got (n=%g h=%g w=%g m8=%g) expected (n=%g h=%g w=%g m8=%g)\n",
             pc.name, r, gn, gh_, gw_, g8, en, eh, ew, e8);
  }
  printf("[probe %s] %s (%d/128 rows differ)\n", pc.name, bad ? "MISMATCH" : "MATCH", bad);
  if (bad) {
    printf("[probe %s] first 40 rows as loaded (n,h,w):", pc.name);
    for (int r = 0; r < 40; ++r) {
      auto elem = [&](int c) {
        const int chunk = (c / 8) ^ (r & 7);
        uint16_t v;
        memcpy(&v, &ho[r * 128 + chunk * 16 + (c % 8) * 2], 2);
        return bf2f(v);
      };
      printf(" (%g,%g,%g)", elem(0), elem(1), elem(2));
    }
    printf("\n");
  }
  cudaFree(dx);
  cudaFree(dout);
  return bad;
}

// ------------------------------------------------------------------ conv check
struct ConvCase {
  const char* name;
  int N, H, W, Cin, Cout, stride;
  int ps;     // pixel-shuffle store
  int act;    // Act
  int stats;  // fused BN statistics
};

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static void fill_fprop_problem(IgemmProblem& p, const ConvCase& cc, const __nv_bfloat16* x,
                               const __nv_bfloat16* w, __nv_bfloat16* y, const float* bias,
                               const float* slope, float* stats) {
  const int OH = (cc.H + 2 - 3) / cc.stride + 1, OW = (cc.W + 2 - 3) / cc.stride + 1;
  p = IgemmProblem{};
  p.x = x; p.NB = cc.N; p.H = cc.H; p.W = cc.W; p.Cin = cc.Cin;
  p.GH = OH; p.GW = OW; p.trav_stride = cc.stride;
  p.lower_w = p.lower_h = -1;
  p.upper_w = p.upper_h = -1;
  p.w = w; p.Cout = cc.Cout; p.Ktot = 9 * cc.Cin; p.num_taps = 9;
  for (int t = 0; t < 9; ++t) {
    p.taps.off_h[t] = t / 3;
    p.taps.off_w[t] = t % 3;
    p.taps.k_off[t] = t * cc.Cin;
  }
  p.out = y;
  if (cc.ps) {
    p.OH = OH * 2; p.OW = OW * 2; p.ldc = cc.Cout / 4; p.ps_c = cc.Cout / 4;
  } else {
    p.OH = OH; p.OW = OW; p.ldc = cc.Cout; p.ps_c = 0;
  }
  p.osy = p.osx = 1; p.opy = p.opx = 0;
  p.bias = bias; p.act = cc.act; p.slope = 0.01f; p.slope_ptr = slope; p.stats = stats; p.stats_rows = igemm_max_ctas();
}

static int run_conv(const ConvCase& cc, bool timing) {
  const int OH = (cc.H + 2 - 3) / cc.stride + 1, OW = (cc.W + 2 - 3) / cc.stride + 1;
  const size_t nx = (size_t)cc.N * cc.H * cc.W * cc.Cin, nw = (size_t)cc.Cout * 9 * cc.Cin,
               ny = (size_t)cc.N * OH * OW * cc.Cout;
  std::vector<uint16_t> hx(nx), hw(nw);
  std::vector<float> hb(cc.Cout);
  for (auto& v : hx) v = f2bf(frand());
  for (auto& v : hw) v = f2bf(frand() * 0.06f);
  for (auto& v : hb) v = frand() * 0.1f;
  __nv_bfloat16 *dx, *dw, *dy, *dref;
  float *dbias, *dslope, *dstats;
  CK(cudaMalloc(&dx, nx * 2)); CK(cudaMalloc(&dw, nw * 2)); CK(cudaMalloc(&dy, ny * 2));
  CK(cudaMalloc(&dref, ny * 2)); CK(cudaMalloc(&dbias, cc.Cout * 4)); CK(cudaMalloc(&dslope, 4));
  const int srows = igemm_max_ctas();
  CK(cudaMalloc(&dstats, (size_t)srows * 2 * cc.Cout * 4));
  CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hb.data(), cc.Cout * 4, cudaMemcpyHostToDevice));
  const float slope = 0.25f;
  CK(cudaMemcpy(dslope, &slope, 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0xFF, ny * 2));
  CK(cudaMemset(dstats, 0xFF, (size_t)srows * 2 * cc.Cout * 4));   // the kernel must overwrite every row

  IgemmProblem p;
  fill_fprop_problem(p, cc, dx, dw, dy, dbias, dslope, cc.stats ? dstats : nullptr);
  if (int rc = igemm_launch(p, 0)) {
    printf("[conv %s] launch failed rc=%d: %s\n", cc.name, rc, igemm_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[conv %s] kernel error: %s\n", cc.name, cudaGetErrorString(e));
    exit(3);
  }
  SimtConv sc{cc.N, cc.H, cc.W, cc.Cin, OH, OW, cc.Cout, 3, 3, cc.stride, 1};
  conv_fprop_simt(sc, dx, dw, dbias, cc.act, 0.01f, dslope, dref, nullptr, 0);
  CK(cudaDeviceSynchronize());
  std::vector<uint16_t> hy(ny), href(ny);
  CK(cudaMemcpy(hy.data(), dy, ny * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(href.data(), dref, ny * 2, cudaMemcpyDeviceToHost));
  double num = 0, den = 0, maxabs = 0;
  size_t nbad = 0;
  std::vector<double> ssum(cc.Cout, 0.0), ssq(cc.Cout, 0.0);
  for (int n = 0; n < cc.N; ++n)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow)
        for (int co = 0; co < cc.Cout; ++co) {
          const float r = bf2f(href[(((size_t)n * OH + oh) * OW + ow) * cc.Cout + co]);
          size_t yi;
          if (cc.ps) {
            const int C4 = cc.Cout / 4, sub = co / C4, ch = co % C4;
            yi = (((size_t)n * OH * 2 + oh * 2 + (sub >> 1)) * OW * 2 + ow * 2 + (sub & 1)) * C4 + ch;
          } else {
            yi = (((size_t)n * OH + oh) * OW + ow) * cc.Cout + co;
          }
          const float g = bf2f(hy[yi]);
          const double d = (double)g - r;
          num += d * d; den += (double)r * r;
          if (fabs(d) > maxabs) maxabs = fabs(d);
          if (!(fabs(d) <= 0.02 + 0.02 * fabs(r))) {
            if (nbad < 8)
              printf("[conv %s] bad n=%d oh=%d ow=%d co=%d got %g ref %g\n", cc.name, n, oh, ow, co,
                     g, r);
            ++nbad;
          }
          ssum[co] += g; ssq[co] += (double)g * g;
        }
  const double rel = sqrt(num / (den + 1e-30));
  int fail = nbad > 0 || !(rel < 1e-2);
  printf("[conv %s] rel_l2=%.3e max_abs=%.3e bad=%zu -> %s\n", cc.name, rel, maxabs, nbad,
         fail ? "FAIL" : "ok");
  if (cc.stats) {
    std::vector<float> hr((size_t)srows * 2 * cc.Cout);
    CK(cudaMemcpy(hr.data(), dstats, hr.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> hs(2 * cc.Cout, 0.0);
    for (int r = 0; r < srows; ++r)
      for (int c = 0; c < 2 * cc.Cout; ++c) hs[c] += hr[(size_t)r * 2 * cc.Cout + c];
    double worst = 0;
    for (int c = 0; c < cc.Cout; ++c) {
      worst = fmax(worst, fabs(hs[c] - ssum[c]) / (fabs(ssum[c]) + 1.0));
      worst = fmax(worst, fabs(hs[cc.Cout + c] - ssq[c]) / (fabs(ssq[c]) + 1.0));
    }
    printf("[conv %s] fused stats worst rel err %.3e -> %s\n", cc.name, worst,
           worst < 1e-3 ? "ok" : "FAIL");
    fail |= !(worst < 1e-3);
  }
  if (timing && getenv("PM_TIMING")) {
    long long* dcnt;
    CK(cudaMalloc(&dcnt, 64));
    CK(cudaMemset(dcnt, 0, 64));
    igemm_set_pm_debug(dcnt);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) igemm_launch(p, 0);
    CK(cudaDeviceSynchronize());
    igemm_set_pm_debug(nullptr);
    long long h[8];
    CK(cudaMemcpy(h, dcnt, 64, cudaMemcpyDeviceToHost));
    if (h[3])
      printf("[phases %s] CTA 0, cycles per tile: MMA warp waits accumulator %lld, box %lld, issues %lld | epilogue warp per M "
             "tile: waits %lld, works %lld | kernel %lld cycles, %lld tiles, %lld M tiles\n", cc.name, h[0] / h[3], h[1] / h[3],
             h[2] / h[3], h[6] ? h[4] / h[6] : 0, h[6] ? h[5] / h[6] : 0, h[7] / reps, h[3] / reps, h[6] / reps);
    cudaFree(dcnt);
  }
  if (timing) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) igemm_launch(p, 0);
    cudaEventRecord(e0);
    const int iters = 50;
    for (int i = 0; i < iters; ++i) igemm_launch(p, 0);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * cc.N * OH * OW * cc.Cout * 9.0 * cc.Cin;
    printf("[time %s] %.2f us/launch  %.1f TFLOP/s (back-to-back, L2-warm)\n", cc.name,
           ms / iters * 1e3, flops / (ms / iters * 1e-3) / 1e12);
  }
  cudaFree(dx); cudaFree(dw); cudaFree(dy); cudaFree(dref); cudaFree(dbias); cudaFree(dslope);
  cudaFree(dstats);
  return fail;
}

int main(int argc, char** argv) {
  int dev_count = 0;
  CK(cudaGetDeviceCount(&dev_count));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount);
  int fails = 0;
  const ProbeCase probes[] = {
      {"s1_pad1_start0_tap00", 3, 12, 12, 64, -1, -1, 1, 12, 12, 0, 0, 0},
      {"s1_pad1_start0_tap22", 3, 12, 12, 64, -1, -1, 1, 12, 12, 0, 2, 2},
      {"s1_pad1_start200_tap12", 3, 12, 12, 64, -1, -1, 1, 12, 12, 200, 1, 2},
      {"s1_pad1_tail_tap11", 3, 12, 12, 64, -1, -1, 1, 12, 12, 384, 1, 1},
      {"s2_pad1_start0_tap00", 4, 12, 12, 64, -1, -1, 2, 6, 6, 0, 0, 0},
      {"s2_pad1_start0_tap21", 4, 12, 12, 64, -1, -1, 2, 6, 6, 0, 2, 1},
      {"dgrad2_lower0_tap11", 4, 6, 6, 64, 0, 0, 1, 6, 6, 0, 1, 1},
      {"dgrad2_lower0_tap01", 4, 6, 6, 64, 0, 0, 1, 6, 6, 20, 0, 1},
  };
  for (const auto& pc : probes) fails += run_probe(pc) ? 1 : 0;

  const ConvCase convs[] = {
      {"c64_24x24_n4", 4, 24, 24, 64, 64, 1, 0, ACT_NONE, 1},
      {"c64_12x12_n3_tail", 3, 12, 12, 64, 64, 1, 0, ACT_PRELU, 1},
      {"c128_12x12_n5", 5, 12, 12, 128, 128, 1, 0, ACT_LEAKY, 1},
      {"c64_256_ps_24x24_n2", 2, 24, 24, 64, 256, 1, 1, ACT_PRELU, 0},
      {"c64_s2_24x24_n4", 4, 24, 24, 64, 64, 2, 0, ACT_NONE, 1},
      {"c256_512_6x6_n8", 8, 6, 6, 256, 512, 1, 0, ACT_RELU, 0},
      {"c128_256_s2_12x12_n8", 8, 12, 12, 128, 256, 2, 0, ACT_NONE, 1},
  };
  if (argc > 1 && argv[1][0] == 'p') {
    // halo-fed kernel for the 64 -> 64 channel layers, pixels on M (igemm_pm.cu): "harness_igemm pm [grp]";
    // PM_TIMING=1 adds the phase counters of CTA 0
    igemm_set_transposed(1);
    igemm_set_pm(1);
    if (argc > 2) igemm_set_pm_grp(atoi(argv[2]));
    printf("-- halo-fed kernel, pixels on M (64 -> 64)\n");
    const ConvCase pm_cases[] = {
        {"p_c64_24x24_n4", 4, 24, 24, 64, 64, 1, 0, ACT_NONE, 1},
        {"p_c64_12x12_n3", 3, 12, 12, 64, 64, 1, 0, ACT_PRELU, 1},
        {"p_c64_9x7_n2_odd", 2, 9, 7, 64, 64, 1, 0, ACT_LEAKY, 1},
        {"p_c64_48x48_n2", 2, 48, 48, 64, 64, 1, 0, ACT_NONE, 1},
        {"p_c64_96x96_n3", 3, 96, 96, 64, 64, 1, 0, ACT_RELU, 0},
        {"p_c64_24x24_n150_multi", 150, 24, 24, 64, 64, 1, 0, ACT_NONE, 1},
        {"p_c64_5x40_n7", 7, 5, 40, 64, 64, 1, 0, ACT_NONE, 1},
    };
    for (const auto& cc : pm_cases) {
      IgemmProblem q;
      fill_fprop_problem(q, cc, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
      if (!igemm_pm_supported(q)) {
        printf("[conv %s] NOT routed to the halo-fed kernel\n", cc.name);
        ++fails;
      }
      fails += run_conv(cc, false);
    }
    const ConvCase pm_big[] = {
        {"trunk_c64_24x24_n64", 64, 24, 24, 64, 64, 1, 0, ACT_NONE, 1},
        {"vgg_c64_96x96_n64", 64, 96, 96, 64, 64, 1, 0, ACT_RELU, 0},
    };
    for (const auto& cc : pm_big) fails += run_conv(cc, true);
    igemm_set_pm(0);
    for (const auto& cc : pm_big) fails += run_conv(cc, true);      // same shapes on the im2col-fed transposed kernel
    printf("harness: %d failure(s)\n", fails);
    return fails ? 1 : 0;
  }
  igemm_set_pm(0);
  igemm_set_transposed(1);
  printf("-- default engine (transposed tiles for Cout <= 128)\n");
  for (const auto& cc : convs) fails += run_conv(cc, false);
  igemm_set_transposed(0);
  printf("-- im2col-fed kernel\n");
  for (const auto& cc : convs) fails += run_conv(cc, false);
  igemm_set_transposed(1);
  if (argc > 1) {
    const ConvCase big[] = {
        {"trunk_c64_24x24_n64", 64, 24, 24, 64, 64, 1, 0, ACT_NONE, 1},
        {"suffix_c64_256_48x48_n64", 64, 48, 48, 64, 256, 1, 1, ACT_PRELU, 0},
        {"vgg_c128_48x48_n64", 64, 48, 48, 128, 128, 1, 0, ACT_RELU, 0},
        {"vgg_c256_24x24_n64", 64, 24, 24, 256, 256, 1, 0, ACT_RELU, 0},
        {"vgg_c512_12x12_n64", 64, 12, 12, 512, 512, 1, 0, ACT_RELU, 0},
    };
    for (const auto& cc : big) fails += run_conv(cc, true);
  }
  printf("harness: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
