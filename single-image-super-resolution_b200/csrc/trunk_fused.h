// Internal C++ interface of the fused persistent trunk forward (trunk_fused.cu; not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace sisr {

constexpr int kTrunkMaxLayers = 40;

// Per-layer tables (host pointers to device memory).  Layer l reads the output `a` of layer l - 1 (layer 0 reads
// x0); `residual_layer`: -1 none, -2 = x0 (the global / first block's skip), l' >= 0 = output of layer l'.
// slope: PReLU slope (one float) or nullptr (no activation).  aux: [4][64] = scale, shift, mean, invstd.
struct TrunkLayerHost {
  const float* bias;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  const float* slope;
  float* aux;
  int residual_layer, reserved;
};

struct TrunkLayerDev {
  const float* bias;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  const float* slope;
  float* aux;
  const __nv_bfloat16* residual;
};

bool trunk_fused_supported(int nb, int h, int w, int n_layers);
size_t trunk_fused_workspace_bytes(int n_layers);
// x0: [nb, h, w, 64] bf16.  weights: prepared bf16 weight matrix, layer l = rows [l * w_row_stride, + 64) of
// 576 columns ([Cout][3][3][Cin]).  y_all / a_all: [n_layers][nb, h, w, 64] bf16 outputs (conv output before BN;
// layer output after BN / activation / residual).  Returns 0 on success.
int trunk_fused_forward(const __nv_bfloat16* x0, int nb, int h, int w, const __nv_bfloat16* weights,
                        int w_row_stride, const TrunkLayerHost* layers, int n_layers, __nv_bfloat16* y_all,
                        __nv_bfloat16* a_all, float momentum, float eps, void* workspace, cudaStream_t stream);
const char* trunk_fused_last_error();

}  // namespace sisr
