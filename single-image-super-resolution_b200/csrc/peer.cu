// SyncBN exchange over NVLink peer memory: one small kernel per exchange instead of a collective
// launch.  Every rank pushes its vector into its row of the slot on ALL ranks (stores through the
// peer mapping), waits until the rows of all ranks carry the current epoch, and adds them in rank
// order - so every rank computes bit-identical sums.  A slot is reused only one training step later;
// the gradient all-reduces in between keep the ranks within one step of each other.
#include "peer.h"
#include "peer_device.cuh"

#include <stdio.h>

namespace sisr {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
peer_allreduce_kernel(const __grid_constant__ PeerTable t, int slot, float* __restrict__ buf, int n) {
  __shared__ float s_vals[kPeerSlotFloats];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_vals[i] = buf[i];
  __syncthreads();
  exchange(t, slot, s_vals, n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = s_vals[i];
}

constexpr int kSyncThreads = 1024;
__global__ void __launch_bounds__(kSyncThreads)
bn_finalize_sync_kernel(const __grid_constant__ PeerTable t, int slot, const float* __restrict__ stats,
                        int stats_rows, float count, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ running_mean,
                        float* __restrict__ running_var, long long* __restrict__ num_batches,
                        float momentum, float eps, float* __restrict__ scale,
                        float* __restrict__ shift, float* __restrict__ mean_out,
                        float* __restrict__ invstd_out, int C) {
  __shared__ float s_vals[kPeerSlotFloats];
  __shared__ float s_part[kSyncThreads];
  // local partial rows (one per persistent conv CTA): thread = (row lane, column), coalesced and
  // independent loads, then a shared-memory sum over the row lanes
  const int n = 2 * C;                       // <= 1024
  const int lanes = kSyncThreads / n;        // >= 1
  const int col = threadIdx.x % n, rl = threadIdx.x / n;
  float acc = 0.f;
  if (rl < lanes) {
#pragma unroll 8
    for (int r = rl; r < stats_rows; r += lanes) acc += stats[static_cast<size_t>(r) * n + col];
  }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < n) {
    for (int l = 1; l < lanes; ++l) acc += s_part[l * n + threadIdx.x];
    s_vals[threadIdx.x] = acc;
  }
  __syncthreads();
  exchange(t, slot, s_vals, n);
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
  return sizeof(uint2) * static_cast<size_t>(kPeerSlots) * kPeerMaxWorld * kPeerSlotFloats +
         sizeof(uint32_t) * static_cast<size_t>(kPeerSlots);
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
  if (!table_ok(t, slot, 2 * C) || 2 * C > kSyncThreads) return 1;
  bn_finalize_sync_kernel<<<1, kSyncThreads, 0, s>>>(t, slot, stats, stats_rows, count, gamma, beta,
                                                 running_mean, running_var, num_batches, momentum, eps,
                                                 scale, shift, mean, invstd, C);
  return check();
}

}  // namespace sisr
