// PSNR and SSIM of image batches (NCHW fp32), the quality metrics the reference lists as a todo
// (README.md:88) for its viewer flow (visualisation.py:46-52: LR / SR = G(LR) / HR / UR = G(HR)).
//   PSNR = 10 log10(L^2 / mean squared error), per image over all channels
//   SSIM = mean over channels and over the valid window positions of the Wang et al. index with an
//          11 x 11 Gaussian window (sigma 1.5), K1 = 0.01, K2 = 0.03, dynamic range L
// One block = one 16 x 16 tile of window positions of one (image, channel) plane: both 26 x 26 input
// patches staged in shared memory, the five Gaussian moments accumulated per thread, block reduction, one
// atomic per block and quantity.  HBM-bound in principle (2 x 4 B per pixel), tiny in practice.
#include "metrics.h"

#include <math.h>

namespace sisr {

namespace {

constexpr int kWin = 11, kTile = 16, kPatch = kTile + kWin - 1;

struct Gauss {
  float w[kWin];
};

__global__ void __launch_bounds__(kTile * kTile)
psnr_ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int H, int W, Gauss g,
                 float c1, float c2, float* __restrict__ sq_err, float* __restrict__ ssim_sum) {
  __shared__ float sa[kPatch][kPatch + 1], sb[kPatch][kPatch + 1];
  __shared__ float red[2][kTile * kTile / 32];
  const int plane = blockIdx.z, n = plane / C;
  const float* pa = a + static_cast<size_t>(plane) * H * W;
  const float* pb = b + static_cast<size_t>(plane) * H * W;
  const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
  const int tid = threadIdx.y * kTile + threadIdx.x;
  for (int i = tid; i < kPatch * kPatch; i += kTile * kTile) {
    const int py = i / kPatch, px = i - py * kPatch;
    const int y = y0 + py, x = x0 + px;
    const bool in = y < H && x < W;
    sa[py][px] = in ? pa[static_cast<size_t>(y) * W + x] : 0.f;
    sb[py][px] = in ? pb[static_cast<size_t>(y) * W + x] : 0.f;
  }
  __syncthreads();
  // squared error: every pixel of the plane belongs to exactly one tile's top-left 16 x 16 corner
  float se = 0.f;
  {
    const int y = y0 + threadIdx.y, x = x0 + threadIdx.x;
    if (y < H && x < W) {
      const float d = sa[threadIdx.y][threadIdx.x] - sb[threadIdx.y][threadIdx.x];
      se = d * d;
    }
  }
  float ss = 0.f;
  const int oy = y0 + threadIdx.y, ox = x0 + threadIdx.x;
  if (oy + kWin <= H && ox + kWin <= W) {
    float ma = 0.f, mb = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int dy = 0; dy < kWin; ++dy) {
      float ra = 0.f, rb = 0.f, raa = 0.f, rbb = 0.f, rab = 0.f;
#pragma unroll
      for (int dx = 0; dx < kWin; ++dx) {
        const float va = sa[threadIdx.y + dy][threadIdx.x + dx], vb = sb[threadIdx.y + dy][threadIdx.x + dx];
        const float w = g.w[dx];
        ra = fmaf(w, va, ra);
        rb = fmaf(w, vb, rb);
        raa = fmaf(w, va * va, raa);
        rbb = fmaf(w, vb * vb, rbb);
        rab = fmaf(w, va * vb, rab);
      }
      const float w = g.w[dy];
      ma = fmaf(w, ra, ma);
      mb = fmaf(w, rb, mb);
      aa = fmaf(w, raa, aa);
      bb = fmaf(w, rbb, bb);
      ab = fmaf(w, rab, ab);
    }
    const float va = aa - ma * ma, vb = bb - mb * mb, cov = ab - ma * mb;
    ss = ((2.f * ma * mb + c1) * (2.f * cov + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = se;
    red[1][tid >> 5] = ss;
  }
  __syncthreads();
  if (tid == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int i = 0; i < kTile * kTile / 32; ++i) {
      t0 += red[0][i];
      t1 += red[1][i];
    }
    atomicAdd(&sq_err[n], t0);
    atomicAdd(&ssim_sum[n], t1);
  }
}

__global__ void psnr_ssim_finish_kernel(const float* sq_err, const float* ssim_sum, int N, float range,
                                        float n_pix, float n_win, float* psnr, float* ssim) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float mse = sq_err[n] / n_pix;
  psnr[n] = mse > 0.f ? 10.f * log10f(range * range / mse) : INFINITY;
  ssim[n] = ssim_sum[n] / n_win;
}

}  // namespace

int psnr_ssim(const float* a, const float* b, int N, int C, int H, int W, float range, float* workspace,
              float* psnr, float* ssim, cudaStream_t s) {
  if (H < kWin || W < kWin || N <= 0 || C <= 0) return 1;
  Gauss g;
  double sum = 0.0;
  for (int i = 0; i < kWin; ++i) {
    const double d = i - (kWin - 1) / 2.0;
    g.w[i] = static_cast<float>(exp(-d * d / (2.0 * 1.5 * 1.5)));
    sum += g.w[i];
  }
  for (int i = 0; i < kWin; ++i) g.w[i] = static_cast<float>(g.w[i] / sum);
  cudaMemsetAsync(workspace, 0, sizeof(float) * 2 * N, s);
  const dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N * C);
  const float c1 = (0.01f * range) * (0.01f * range), c2 = (0.03f * range) * (0.03f * range);
  psnr_ssim_kernel<<<grid, dim3(kTile, kTile), 0, s>>>(a, b, C, H, W, g, c1, c2, workspace, workspace + N);
  psnr_ssim_finish_kernel<<<(N + 127) / 128, 128, 0, s>>>(workspace, workspace + N, N, range,
                                                          static_cast<float>(C) * H * W,
                                                          static_cast<float>(C) * (H - kWin + 1) * (W - kWin + 1),
                                                          psnr, ssim);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

}  // namespace sisr
