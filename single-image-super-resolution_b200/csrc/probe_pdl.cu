// Stand-alone probe (not part of the library): what does programmatic dependent launch buy for a chain of
// short dependent kernels replayed from a CUDA graph?  Chain of kLen kernels, each reading the previous
// kernel's output (a 4.7 MB tensor, the generator-trunk size), timed as (a) plain launches, (b) launches with
// the programmatic-stream-serialization attribute + griddepcontrol.wait / launch_dependents in the kernel.
// A "prologue" of `spin` clock cycles before the wait stands in for barrier init / TMEM alloc / descriptor
// prefetch of the conv kernels.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void step_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n, int pdl, int spin) {
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { }
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 v = in[i];
    v.x += 1u; v.y ^= v.x; v.z += v.y; v.w ^= v.z;
    out[i] = v;
  }
}

static float run(cudaStream_t s, uint4* a, uint4* b, long long n, int len, int pdl, int spin, int blocks, int smem) {
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < len; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, step_kernel, (const uint4*)(i & 1 ? b : a), (i & 1 ? a : b), n, pdl, spin));
  }
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / (10 * len);
}

int main() {
  const long long n = 36864LL * 64 * 2 / 16;      // 4.7 MB of bf16 as uint4
  uint4 *a, *b;
  CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
  CK(cudaMemset(a, 0, n * 16)); CK(cudaMemset(b, 0, n * 16));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  CK(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int len = 200;
  for (int smem : {0, 100 * 1024, 200 * 1024})
    for (int blocks : {148, 592})
      for (int spin : {0, 2000, 6000}) {
        if (smem && blocks != 148) continue;
        const float t0 = run(s, a, b, n, len, 0, spin, blocks, smem);
        const float t1 = run(s, a, b, n, len, 1, spin, blocks, smem);
        printf("smem %3d KB  blocks %4d  prologue %5d clk:  plain %6.2f us/kernel   pdl %6.2f us/kernel   saved %5.2f us\n",
               smem / 1024, blocks, spin, t0, t1, t0 - t1);
      }
  // verify the chain result (every element incremented len * 13 times on .x)
  uint4 h; CK(cudaMemcpy(&h, a, 16, cudaMemcpyDeviceToHost));
  printf("check x = %u\n", h.x);
  return 0;
}
