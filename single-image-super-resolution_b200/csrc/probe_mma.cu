// Hardware probe (B200), stand-alone (not part of the library): how many cycles does ONE tcgen05.mma
// (bf16, K = 16) take as a function of M, N, the operand sources and the CTA group?
//   SS  : A and B through shared-memory descriptors (what the conv kernels use)
//   TS  : A from tensor memory (kind::f16 [d], [a_tmem], b_desc), B from shared memory
//   2CTA: cta_group::2, M = 256 (128 rows per CTA), each CTA holds half of the B rows
// The conv kernels' per-layer numbers fit "cycles ~ max(A rows, B rows) read from shared memory per K = 16
// step" (profiles/r1_notes.md); this probe measures that law directly and tells whether the 64-channel
// layers (25 % of peak in the SS form) would gain from A-in-TMEM and how much the CTA pair buys.
// Operand contents are irrelevant (zero-filled shared memory); R back-to-back instructions accumulate into
// the same TMEM columns, one commit, one mbarrier wait, clock64 around the whole chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_mma csrc/probe_mma.cu && timeout 60 build/probe_mma
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace sisr;

// (the cta_group::2 helpers live in ptx.cuh: the pixels-on-M conv kernel uses them)

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int kReps = 512;

// mode 0: SS, 1: TS.  One CTA, 128 threads.  A tile: 128 rows x 64 k (16 KB), B tile: N rows x 64 k.
__global__ void __launch_bounds__(128)
probe_1cta(int M, int N, int mode, long long* cycles, int nacc = 1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
    const uint32_t tmem_a = tmem + 256;           // columns 256.. hold the A operand in the TS form
    const long long t0 = clock64();
    for (int r = 0; r < kReps; ++r) {
      const int k = r & 3;
      const uint64_t db = umma_smem_desc(b_addr + k * 32, 16, 1024);
      // nacc > 1: consecutive instructions go to different accumulators (is the floor a dependency latency?)
      const uint32_t d_acc = tmem + (r % nacc) * N;
      if (mode == 2) {
        // both operands MN-major (the weight-gradient form: K = pixels): [16 pixel rows][64 channels] tiles of
        // 128 B rows, 64-channel groups 8 KB apart (LBO), a K step = 2048 B
        const uint32_t idesc_mn = umma_idesc_bf16(M, N, 1, 1);
        umma_bf16(d_acc, umma_smem_desc(a_addr + k * 2048, 8192, 1024), umma_smem_desc(b_addr + k * 2048, 8192, 1024),
                  idesc_mn, r >= nacc);
      } else if (mode == 0)
        umma_bf16(d_acc, umma_smem_desc(a_addr + k * 32, 16, 1024), db, idesc, r >= nacc);
      else
        umma_bf16_ts(d_acc, tmem_a + k * 8, db, idesc, r >= nacc);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    *cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// CTA pair: M = 256, each CTA holds 128 A rows and N/2 B rows.  Leader issues, both wait.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe_2cta(int N, long long* cycles, int nacc = 1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc_pair(smem_u32(&slot), 512);
    tmem_relinquish_pair();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
    const long long t0 = clock64();
    for (int r = 0; r < kReps; ++r) {
      const int k = r & 3;
      umma_bf16_pair(tmem + (r % nacc) * N, umma_smem_desc(a_addr + k * 32, 16, 1024),
                     umma_smem_desc(b_addr + k * 32, 16, 1024), idesc, r >= nacc);
    }
    umma_commit_pair(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    *cycles = t1 - t0;
  }
  if (rank == 1 && warp == 1 && lane == 0) mbar_wait(smem_u32(&bar), 0);     // multicast commit lands here too
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 8));
  const int smem = 16384 + 32768 + 1024;
  CK(cudaFuncSetAttribute(probe_1cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  auto report = [&](const char* what, int M, int N) {
    long long c;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    const double per = static_cast<double>(c) / kReps;
    const double mac = static_cast<double>(M > 128 ? 128 : M) * N * 16 / per;       // per SM
    printf("%-6s M=%3d N=%3d : %7.1f cycles / K=16 instruction   %6.0f MAC/clk/SM (%4.1f %% of 4096)\n", what, M, N,
           per, mac, 100.0 * mac / 4096.0);
  };
  for (int pass = 0; pass < 3; ++pass) {      // the first pass runs on a GPU that has just left its idle state
  printf("pass %d\n", pass);
  for (int mode = 0; mode < 2; ++mode)
    for (int M : {64, 128})
      for (int N : {64, 128, 256}) {
        if (mode == 1 && M == 64) continue;
        for (int rep = 0; rep < 2; ++rep) probe_1cta<<<1, 128, smem>>>(M, N, mode, d);
        report(mode ? "TS" : "SS", M, N);
      }
  for (int N : {64, 128, 256}) {
    for (int rep = 0; rep < 2; ++rep) probe_2cta<<<2, 128, smem>>>(N, d);
    report("2CTA", 256, N);
  }
  for (int N : {64, 128, 256}) {
    for (int rep = 0; rep < 2; ++rep) probe_1cta<<<1, 128, smem>>>(128, N, 2, d);
    report("MN", 128, N);
  }
  // independent accumulators, round robin (SS form; the TS form keeps its A operand in columns 256..)
  for (int nacc : {2, 4})
    for (int N : {64, 128}) {
      if (nacc * N > 256) continue;
      for (int rep = 0; rep < 2; ++rep) probe_1cta<<<1, 128, smem>>>(128, N, 0, d, nacc);
      char what[16];
      snprintf(what, sizeof what, "SSx%d", nacc);
      report(what, 128, N);
      for (int rep = 0; rep < 2; ++rep) probe_2cta<<<2, 128, smem>>>(N, d, nacc);
      snprintf(what, sizeof what, "2Cx%d", nacc);
      report(what, 256, N);
    }
  }
  CK(cudaFree(d));
  return 0;
}
