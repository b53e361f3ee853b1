// Hardware probe (B200): does a tcgen05.mma shared-memory descriptor whose start address is shifted
// by s x 128 B (s pixel rows, not a multiple of the 8-row / 1024 B swizzle atom) read the rows
// s .. s+127 of a 128 B-swizzled tile that TMA wrote at a 1024 B-aligned base?  If it does, the nine
// filter taps of a 3x3 conv can be nine descriptor views of ONE halo tile instead of nine loads.
//   mode 0: descriptor base_offset field = 0          mode 1: base_offset = s & 7
//   K-major A (fprop / dgrad) and MN-major A (wgrad, shift along K) are probed separately.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ptx.cuh"
#include "tmap.h"

using namespace sisr;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7FFF + ((u >> 16) & 1)) >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

__device__ __forceinline__ uint64_t desc_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
  return umma_smem_desc(addr, lbo, sbo) | (static_cast<uint64_t>(bo & 7) << 49);
}

constexpr int kRows = 256;

// mn_major = 0: D[i][n] = sum_k X0[s+i][k] * W[n][k]          (A K-major rows s..s+127 of X0)
// mn_major = 1: D[m][n] = sum_{k<64} X{m/64}[s+k][m%64] * W[k][n]  (A, B MN-major; X1 = second tile)
__global__ void __launch_bounds__(128)
shift_probe_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                   const __grid_constant__ CUtensorMap tm_w, int s, int mode, int mn_major,
                   float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* x0 = smem;                    // 256 rows x 128 B
  uint8_t* x1 = smem + kRows * 128;      // second 64-channel group (MN-major probe)
  uint8_t* w = smem + 2 * kRows * 128;   // 64 rows x 128 B
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&mma_bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar), 2 * kRows * 128 + 64 * 128);
    tma_load_2d(smem_u32(x0), &tm_x0, smem_u32(&bar), 0, 0);
    tma_load_2d(smem_u32(x1), &tm_x1, smem_u32(&bar), 0, 0);
    tma_load_2d(smem_u32(w), &tm_w, smem_u32(&bar), 0, 0);
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t bo = mode ? static_cast<uint32_t>(s & 7) : 0u;
    if (!mn_major) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = desc_bo(smem_u32(x0) + s * 128 + k * 32, 16, 1024, bo);
        const uint64_t db = umma_smem_desc(smem_u32(w) + k * 32, 16, 1024);
        umma_bf16(tmem, da, db, idesc, k > 0);
      }
    } else {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      for (int j = 0; j < 4; ++j) {
        const uint64_t da = desc_bo(smem_u32(x0) + s * 128 + j * 2048, kRows * 128, 1024, bo);
        const uint64_t db = umma_smem_desc(smem_u32(w) + j * 2048, 8192, 1024);
        umma_bf16(tmem, da, db, idesc, j > 0);
      }
    }
    umma_commit(smem_u32(&mma_bar));
  }
  mbar_wait(smem_u32(&mma_bar), 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t raw[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, raw);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c * 32 + i] = __uint_as_float(raw[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

int main() {
  std::vector<uint16_t> hx0(kRows * 64), hx1(kRows * 64), hw(64 * 64);
  srand(7);
  for (auto& v : hx0) v = f2bf((float)(rand() % 9 - 4));
  for (auto& v : hx1) v = f2bf((float)(rand() % 9 - 4));
  for (auto& v : hw) v = f2bf((float)(rand() % 5 - 2));
  uint16_t *dx0, *dx1, *dw;
  float* dout;
  CK(cudaMalloc(&dx0, hx0.size() * 2));
  CK(cudaMalloc(&dx1, hx1.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMemcpy(dx0, hx0.data(), hx0.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx1, hx1.data(), hx1.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tx0, tx1, tw;
  if (make_tmap_2d_bf16(&tx0, dx0, kRows, 64, 64, 64, kRows) ||
      make_tmap_2d_bf16(&tx1, dx1, kRows, 64, 64, 64, kRows) ||
      make_tmap_2d_bf16(&tw, dw, 64, 64, 64, 64, 64)) {
    printf("tensor map error: %s\n", tmap_last_error());
    return 2;
  }
  const int smem_bytes = 2 * kRows * 128 + 64 * 128 + 1024;
  CK(cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int shifts[] = {0, 8, 1, 2, 3, 5, 7, 26, 27, 28, 53, 54};
  int all_ok[2][2] = {{1, 1}, {1, 1}};
  for (int mn = 0; mn < 2; ++mn)
    for (int mode = 0; mode < 2; ++mode)
      for (int s : shifts) {
        CK(cudaMemset(dout, 0, 128 * 64 * 4));
        shift_probe_kernel<<<1, 128, smem_bytes>>>(tx0, tx1, tw, s, mode, mn, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("[shift %s mode %d s=%d] kernel error: %s\n", mn ? "MN" : "K", mode, s,
                 cudaGetErrorString(e));
          return 3;
        }
        std::vector<float> ho(128 * 64);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) {
              if (!mn)
                ref += bf2f(hx0[(s + m) * 64 + k]) * bf2f(hw[n * 64 + k]);
              else
                ref += bf2f((m < 64 ? hx0 : hx1)[(s + k) * 64 + (m & 63)]) * bf2f(hw[k * 64 + n]);
            }
            if (fabs(ref - ho[m * 64 + n]) > 1e-3) ++bad;
          }
        printf("[shift %s-major base_offset=%s s=%2d] %s (%d/8192 differ)\n", mn ? "MN" : "K ",
               mode ? "s&7" : "0  ", s, bad ? "MISMATCH" : "MATCH", bad);
        if (bad) all_ok[mn][mode] = 0;
      }
  for (int mn = 0; mn < 2; ++mn)
    for (int mode = 0; mode < 2; ++mode)
      printf("SUMMARY %s-major base_offset=%s : %s\n", mn ? "MN" : "K", mode ? "s&7" : "0",
             all_ok[mn][mode] ? "ALL SHIFTS MATCH" : "some shifts mismatch");
  return 0;
}
