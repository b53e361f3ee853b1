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

// cta_group::2 helpers: only this probe uses them (the CTA-pair conv kernel was measured slower and removed)
namespace sisr {
// ---------------------------------------------------------------- CTA pair (cta_group::2)
// Two CTAs of a (2,1,1) cluster on the SMs of one TPC execute ONE tcgen05.mma with M = 256: each CTA
// supplies 128 rows of A and HALF of the B rows from its own shared memory, so the per-SM operand read
// per instruction halves for B.  Only the leader (cluster rank 0) issues the MMA; both CTAs run TMA and
// an epilogue for their own 128 accumulator lanes.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in the leader CTA (peer bit cleared)
__device__ __forceinline__ uint32_t leader_addr(uint32_t smem_addr) { return smem_addr & 0xFEFFFFFFu; }
// arrive on an mbarrier that lives in the LEADER CTA's shared memory (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads issued by either CTA of the pair; the transaction bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c, int w,
                                                        int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_addr(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
      "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the mbarrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}

}  // namespace sisr

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
      if (mode == 0)
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
  CK(cudaFree(d));
  return 0;
}
