// Internal C++ interface: spectral norm + weight layout preparation (see spectral.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace sisr {

// Per-layer descriptors of the batched (one launch chain per network) variants; same layout as
// sisr_sn_layer / sisr_prep_layer of the C ABI.
struct SnLayer {
  const float* w;      // [cout, K] fp32 master weight
  float* u;            // [cout]  module buffer (updated in training)
  float* v;            // [K]
  float* t;            // [K]     scratch
  float* s;            // [cout]  scratch
  float* sigma;        // [1]
  float* u_saved;      // [cout]  copy used by the backward pass of this forward call
  float* v_saved;      // [K]
  int cout, K, training, pad_;
};
struct PrepLayer {
  const float* w;
  const float* sigma;      // nullable
  const float* bias;       // nullable unless bias_perm
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;       // nullable
  float* bias_perm;        // nullable
  int cout, cin, k, ps_r;
};

// Tables travel as kernel parameters (no device-side table, CUDA-graph safe).
constexpr int kMaxBatch = 40;
struct SnTable {
  SnLayer L[kMaxBatch];
  int wtu_begin[kMaxBatch + 1];   // first 128-column block of each layer (training layers only)
  int row_begin[kMaxBatch + 1];   // first weight row of each layer
  int n;
};
struct PrepTable {
  PrepLayer L[kMaxBatch];
  int blk_begin[kMaxBatch + 1];   // first 1024-element block of each layer
  int n;
};

size_t sn_workspace_floats(int Cout, int K);
// `layers` are HOST arrays; one launch chain per <= kMaxBatch layers
int sn_power_iteration_batched(const SnLayer* layers, int n_layers, float eps, cudaStream_t s);
int weight_prep_batched(const PrepLayer* layers, int n_layers, cudaStream_t s);
// w: [Cout, K] fp32 (K = Cin*kh*kw, native flatten order).  training: one power iteration,
// u/v updated in place, sigma written.  eval: sigma = u^T W v with the stored vectors.
int sn_power_iteration(const float* w, float* u, float* v, float* sigma, int Cout, int K,
                       int training, float eps, float* ws, cudaStream_t s);
// sigma / bias / wd / bias_perm may be null.  ps_r > 1 permutes rows so that GEMM column
// (i*r + j)*C/r^2 + c holds PixelShuffle output channel c at sub-pixel (i, j).
int weight_prep(const float* w, const float* sigma, const float* bias, __nv_bfloat16* wf,
                __nv_bfloat16* wd, float* bias_perm, int Cout, int Cin, int KH, int KW, int ps_r,
                cudaStream_t s);
// gp: gradient w.r.t. the prepared weight, fp32 [Cout', kh, kw, Cin]; dw: [Cout, Cin, kh, kw].
int weight_grad_finish(const float* gp, const float* w, const float* u, const float* v,
                       const float* sigma, float* dw, const float* dbias_perm, float* dbias, int Cout,
                       int Cin, int KH, int KW, int ps_r, int accumulate, float* ws, cudaStream_t s);

// Fused split-K reduce + finish: partials [splits][Cout'][kh][kw][Cin] fp32 -> dw [Cout][Cin][kh][kw]
// (one cooperative kernel when spectral norm is on).  dot: one float of scratch.
int weight_grad_reduce_finish(const float* partials, int splits, const float* w, const float* u,
                              const float* v, const float* sigma, float* dw, const float* dbias_perm,
                              float* dbias, int Cout, int Cin, int KH, int KW, int ps_r, int accumulate,
                              float* dot, cudaStream_t s);
void weight_grad_disable_cooperative(int off);   // 1: two ordinary launches instead (debug / fallback)

}  // namespace sisr
