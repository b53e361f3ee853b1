// Device side of the NVLink peer-memory exchange (see peer.cu); shared with the BatchNorm backward
// reduction, whose last CTA runs the exchange itself.
#pragma once
#include <stdio.h>

#include "peer.h"

namespace sisr {

// Low-latency exchange: every value travels as one 8-byte {value, epoch} store (a single NVLink
// transaction, never torn), so the receiver needs no separate flag and the sender no system fence -
// it polls the payload words until they carry the current epoch (same idea as NCCL's LL protocol).
__device__ __forceinline__ uint2* slot_row(void* base, int slot, int r) {
  return reinterpret_cast<uint2*>(base) +
         (static_cast<size_t>(slot) * kPeerMaxWorld + r) * kPeerSlotFloats;
}
__device__ __forceinline__ uint32_t* epoch_ptr(void* base, int slot) {
  uint8_t* p = reinterpret_cast<uint8_t*>(base) +
               sizeof(uint2) * static_cast<size_t>(kPeerSlots) * kPeerMaxWorld * kPeerSlotFloats;
  return reinterpret_cast<uint32_t*>(p) + slot;
}
__device__ __forceinline__ void st_ll(uint2* p, float v, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

// vals: this rank's vector in shared memory (n floats); on return it holds the sum over ranks
// (added in rank order: bit-identical on every rank).
__device__ inline void exchange(const PeerTable& t, int slot, float* vals, int n) {
  __shared__ uint32_t s_epoch;
  void* mine = t.base[t.rank];
  if (threadIdx.x == 0) {
    uint32_t* e = epoch_ptr(mine, slot);
    uint32_t next = *e + 1;
    if (next == 0) next = 1;      // 0 is the state of the freshly zeroed workspace
    *e = next;
    s_epoch = next;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  for (int r = 0; r < t.world; ++r) {
    uint2* dst = slot_row(t.base[r], slot, t.rank);
    for (int i = threadIdx.x; i < n; i += blockDim.x) st_ll(dst + i, vals[i], epoch);
  }
  __syncthreads();    // every thread has read vals[] before it is overwritten below
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < t.world; ++r) {
      const uint2* src = slot_row(mine, slot, r) + i;
      uint2 v = ld_ll(src);
      while (v.y != epoch) {
        if (clock64() - t0 > 20000000000LL) {   // ~10 s: a peer never arrived
          printf("sisr: SyncBN peer exchange timeout (rank %d waits for rank %d, slot %d, epoch %u)\n",
                 t.rank, r, slot, epoch);
          __trap();
        }
        v = ld_ll(src);
      }
      acc += __uint_as_float(v.x);
    }
    vals[i] = acc;
  }
  __syncthreads();
}


}  // namespace sisr
