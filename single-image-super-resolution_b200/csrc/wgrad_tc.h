// Internal C++ interface of the tcgen05 weight-gradient engine (see wgrad_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace sisr {

bool wgrad_tc_supported(int n, int h, int w, int cin, int oh, int ow, int cout, int k, int stride,
                        int pad, int ps_r);
size_t wgrad_tc_workspace_bytes(int n, int h, int w, int cin, int oh, int ow, int cout, int k,
                                int stride, int pad, int ps_r);
// g: fp32 [cout', 3, 3, cin] overwritten; dbias: fp32 [cout'] or null.  g == nullptr: the split-K
// partials [splits][cout'][3][3][cin] are left in `workspace` (wgrad_tc_workspace_bytes) for the caller
// to reduce; *splits_out receives their number.
int wgrad_tc_launch(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* g, float* dbias,
                    void* workspace, int n, int h, int w, int cin, int oh, int ow, int cout, int stride,
                    int ps_r, cudaStream_t s, int* splits_out = nullptr);
const char* wgrad_tc_last_error();
// tools only: 8 device counters that CTA 0 adds its phase cycles to ({producer waits stage, MMA warp waits operands,
// MMA warp issues, k-blocks, epilogue, kernel cycles, launches, grid}); nullptr (default) = no instrumentation
void wgrad_tc_set_debug(long long* counters);

}  // namespace sisr
