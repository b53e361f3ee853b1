// Cross-GPU exchange of small per-layer statistic vectors through NVLink peer memory (SyncBN).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sisr {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerSlots = 512;          // exchanges per step (one per BatchNorm forward / backward)
constexpr int kPeerSlotFloats = 1152;    // >= 2*512 + 1 values per exchange

// Layout of every rank's workspace (identical on all ranks, allocated with cudaMalloc and mapped
// into the peers through CUDA IPC):
//   data  [kPeerSlots][kPeerMaxWorld][kPeerSlotFloats] {fp32 value, u32 epoch}   row r of a slot is
//         written by rank r, 8 bytes per element in one store
//   epoch [kPeerSlots] u32                                    local call counter of the slot
size_t peer_workspace_bytes();

struct PeerTable {
  void* base[kPeerMaxWorld];   // workspace of every rank as mapped in THIS process (own entry included)
  int rank, world;
};

// buf[i] = sum over ranks of buf[i], i < n (n <= kPeerSlotFloats); identical bits on every rank
int peer_allreduce(const PeerTable& t, int slot, float* buf, int n, cudaStream_t s);
// BatchNorm finalize on the GLOBAL batch: adds the local partial rows, exchanges the [2C] sums
// with the peers, then does what bn_finalize does (count = global element count per channel).
int bn_finalize_sync(const PeerTable& t, int slot, const float* stats, int stats_rows, float count,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     long long* num_batches, float momentum, float eps, float* scale, float* shift,
                     float* mean, float* invstd, int C, cudaStream_t s);

}  // namespace sisr
