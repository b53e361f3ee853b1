// SyncBN exchange over NVLink peer memory: one small kernel per exchange instead of a collective
// launch.  Every rank pushes its vector into its row of the slot on ALL ranks (plain stores through
// the peer mapping), publishes an epoch flag with release semantics at system scope, waits until
// the flags of all ranks reached the same epoch, and adds the rows in rank order - so every rank
// computes bit-identical sums.  A slot is reused only one training step later; the gradient
// all-reduces in between keep the ranks within one step of each other.
#include "peer.h"

#include <stdio.h>

namespace sisr {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float* slot_row(void* base, int slot, int r) {
  return reinterpret_cast<float*>(base) +
         (static_cast<size_t>(slot) * kPeerMaxWorld + r) * kPeerSlotFloats;
}
__device__ __forceinline__ uint32_t* flag_ptr(void* base, int slot, int r) {
  uint8_t* p = reinterpret_cast<uint8_t*>(base) +
               sizeof(float) * static_cast<size_t>(kPeerSlots) * kPeerMaxWorld * kPeerSlotFloats;
  return reinterpret_cast<uint32_t*>(p) + static_cast<size_t>(slot) * kPeerMaxWorld + r;
}
__device__ __forceinline__ uint32_t* epoch_ptr(void* base, int slot) {
  return flag_ptr(base, kPeerSlots, 0) + slot;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// vals: this rank's vector in shared memory (n floats); on return it holds the sum over ranks.
__device__ void exchange(const PeerTable& t, int slot, float* vals, int n) {
  __shared__ uint32_t s_epoch;
  void* mine = t.base[t.rank];
  if (threadIdx.x == 0) {
    uint32_t* e = epoch_ptr(mine, slot);
    s_epoch = *e + 1;
    *e = s_epoch;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  for (int r = 0; r < t.world; ++r) {
    float* dst = slot_row(t.base[r], slot, t.rank);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < t.world) {
    st_release_sys(flag_ptr(t.base[threadIdx.x], slot, t.rank), epoch);
    const uint32_t* f = flag_ptr(mine, slot, threadIdx.x);
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) {   // ~10 s: a peer never arrived
        printf("sisr: SyncBN peer exchange timeout (rank %d waits for rank %d, slot %d, epoch %u)\n",
               t.rank, threadIdx.x, slot, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < t.world; ++r) acc += slot_row(mine, slot, r)[i];
    vals[i] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
peer_allreduce_kernel(const __grid_constant__ PeerTable t, int slot, float* __restrict__ buf, int n) {
  __shared__ float s_vals[kPeerSlotFloats];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_vals[i] = buf[i];
  __syncthreads();
  exchange(t, slot, s_vals, n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = s_vals[i];
}

__global__ void __launch_bounds__(kThreads)
bn_finalize_sync_kernel(const __grid_constant__ PeerTable t, int slot, const float* __restrict__ stats,
                        int stats_rows, float count, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ running_mean,
                        float* __restrict__ running_var, long long* __restrict__ num_batches,
                        float momentum, float eps, float* __restrict__ scale,
                        float* __restrict__ shift, float* __restrict__ mean_out,
                        float* __restrict__ invstd_out, int C) {
  __shared__ float s_vals[kPeerSlotFloats];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
#pragma unroll 4
    for (int r = 0; r < stats_rows; ++r) acc += stats[static_cast<size_t>(r) * 2 * C + i];
    s_vals[i] = acc;
  }
  __syncthreads();
  exchange(t, slot, s_vals, 2 * C);
  if (threadIdx.x == 0 && num_batches) *num_batches += 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = s_vals[c] / count;
    const float var = fmaxf(s_vals[C + c] / count - mean * mean, 0.f);
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    const float invstd = rsqrtf(var + eps);
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    mean_out[c] = mean;
    invstd_out[c] = invstd;
  }
}

int check() { return cudaGetLastError() == cudaSuccess ? 0 : 4; }
bool table_ok(const PeerTable& t, int slot, int n) {
  if (t.world < 1 || t.world > kPeerMaxWorld || t.rank < 0 || t.rank >= t.world) return false;
  if (slot < 0 || slot >= kPeerSlots || n < 0 || n > kPeerSlotFloats) return false;
  for (int r = 0; r < t.world; ++r)
    if (!t.base[r]) return false;
  return true;
}

}  // namespace

size_t peer_workspace_bytes() {
  return sizeof(float) * static_cast<size_t>(kPeerSlots) * kPeerMaxWorld * kPeerSlotFloats +
         sizeof(uint32_t) * (static_cast<size_t>(kPeerSlots) * kPeerMaxWorld + kPeerSlots);
}

int peer_allreduce(const PeerTable& t, int slot, float* buf, int n, cudaStream_t s) {
  if (!table_ok(t, slot, n)) return 1;
  peer_allreduce_kernel<<<1, kThreads, 0, s>>>(t, slot, buf, n);
  return check();
}

int bn_finalize_sync(const PeerTable& t, int slot, const float* stats, int stats_rows, float count,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     long long* num_batches, float momentum, float eps, float* scale, float* shift,
                     float* mean, float* invstd, int C, cudaStream_t s) {
  if (!table_ok(t, slot, 2 * C)) return 1;
  bn_finalize_sync_kernel<<<1, kThreads, 0, s>>>(t, slot, stats, stats_rows, count, gamma, beta,
                                                 running_mean, running_var, num_batches, momentum, eps,
                                                 scale, shift, mean, invstd, C);
  return check();
}

}  // namespace sisr
