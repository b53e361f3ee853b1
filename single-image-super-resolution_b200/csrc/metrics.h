// Internal C++ interface of the image-quality metrics (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace sisr {

// a, b: [N, C, H, W] fp32; workspace: 2 * N floats; psnr, ssim: [N].  range = dynamic range L (2 for
// images in [-1, 1]).  Returns 0 on success, 1 for images smaller than the 11 x 11 window.
int psnr_ssim(const float* a, const float* b, int N, int C, int H, int W, float range, float* workspace,
              float* psnr, float* ssim, cudaStream_t s);

}  // namespace sisr
