// extern "C" surface declared in include/sisr_b200.h: argument checking, dispatch between the
// tcgen05 implicit-GEMM engine and the CUDA-core kernels, error reporting.
#include "../../include/sisr_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "conv_simt.h"
#include "conv_thin.h"
#include "elementwise.h"
#include "igemm.h"
#include "linear.h"
#include "metrics.h"
#include "optim.h"
#include "peer.h"
#include "spectral.h"
#include "wgrad_tc.h"

using namespace sisr;

namespace {

thread_local char g_err[512] = "";

int fail(int rc, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return rc;
}
int wrap(int rc, const char* what) {
  if (rc == 0) return 0;
  cudaError_t e = cudaGetLastError();
  return fail(rc, "%s failed (rc=%d, cuda: %s)", what, rc, cudaGetErrorString(e));
}
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline const __nv_bfloat16* B(const sisr_bf16* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }
inline __nv_bfloat16* B(sisr_bf16* p) { return reinterpret_cast<__nv_bfloat16*>(p); }

bool desc_ok(const sisr_conv_desc* d) {
  if (!d || d->n <= 0 || d->h <= 0 || d->w <= 0 || d->cin <= 0 || d->cout <= 0 || d->k <= 0) return false;
  if (d->oh != (d->h + 2 * d->pad - d->k) / d->stride + 1) return false;
  if (d->ow != (d->w + 2 * d->pad - d->k) / d->stride + 1) return false;
  if (d->ps_r != 0 && d->ps_r != 1 && d->ps_r != 2) return false;
  return true;
}
bool tc_shape(const sisr_conv_desc* d) {
  return d->k == 3 && d->pad == 1 && (d->stride == 1 || d->stride == 2) && d->cin % 64 == 0 &&
         d->cout % 64 == 0 && d->cin <= 512 && d->cout <= 512 && (d->ps_r < 2 || (d->stride == 1 && (d->cout / 4) % 32 == 0));
}
// thin (3-channel) edge layers: stride 1, same-size
bool thin_geometry(const sisr_conv_desc* d) {
  return d->stride == 1 && d->oh == d->h && d->ow == d->w && d->ps_r < 2 && 2 * d->pad == d->k - 1;
}
ThinConv thin_of(const sisr_conv_desc* d, int cs, int cw) {
  return ThinConv{d->n, d->h, d->w, cs, cw, d->k, d->pad};
}
SimtConv simt_of(const sisr_conv_desc* d) {
  return SimtConv{d->n, d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->k, d->k, d->stride, d->pad};
}

}  // namespace

extern "C" {

const char* sisr_last_error(void) { return g_err; }
int sisr_abi_version(void) { return 2; }
int sisr_stats_rows(void) { return igemm_max_ctas(); }
int sisr_debug_transposed(int on) { igemm_set_transposed(on); return 0; }
int sisr_debug_pm_mode(int on) { igemm_set_pm(on); return 0; }
int sisr_conv_uses_tensor_cores(const sisr_conv_desc* d) { return desc_ok(d) && tc_shape(d) ? 1 : 0; }

// ------------------------------------------------------------------ layout
int sisr_nchw_f32_to_nhwc_bf16(const float* x, sisr_bf16* y, int n, int c, int h, int w, void* s) {
  return wrap(nchw_f32_to_nhwc_bf16(x, B(y), n, c, h, w, S(s)), "nchw_f32_to_nhwc_bf16");
}
int sisr_nhwc_bf16_to_nchw_f32(const sisr_bf16* x, float* y, int n, int c, int h, int w, void* s) {
  return wrap(nhwc_bf16_to_nchw_f32(B(x), y, n, c, h, w, S(s)), "nhwc_bf16_to_nchw_f32");
}
int sisr_transpose_bf16(const sisr_bf16* x, sisr_bf16* y, int batch, int rows, int cols, void* s) {
  return wrap(transpose_bf16(B(x), B(y), batch, rows, cols, S(s)), "transpose_bf16");
}
int sisr_tanh_bwd_nchw_to_nhwc(const float* dout, const float* y, sisr_bf16* dpre, int n, int c, int h,
                               int w, void* s) {
  return wrap(tanh_bwd_nchw_to_nhwc(dout, y, B(dpre), n, c, h, w, S(s)), "tanh_bwd_nchw_to_nhwc");
}

// ------------------------------------------------------------------ spectral norm / weights
size_t sisr_sn_workspace_floats(int cout, int k) { return sn_workspace_floats(cout, k); }
int sisr_sn_power_iteration(const float* w, float* u, float* v, float* sigma, int cout, int k,
                            int training, float eps, float* ws, void* s) {
  return wrap(sn_power_iteration(w, u, v, sigma, cout, k, training, eps, ws, S(s)),
              "sn_power_iteration");
}
static_assert(sizeof(sisr_sn_layer) == sizeof(SnLayer), "sisr_sn_layer layout");
static_assert(sizeof(sisr_prep_layer) == sizeof(PrepLayer), "sisr_prep_layer layout");
int sisr_sn_power_iteration_batched(const sisr_sn_layer* layers, int n_layers, float eps, void* s) {
  if (n_layers < 0 || (n_layers > 0 && !layers)) return fail(1, "sn_power_iteration_batched: bad arguments");
  return wrap(sn_power_iteration_batched(reinterpret_cast<const SnLayer*>(layers), n_layers, eps, S(s)),
              "sn_power_iteration_batched");
}
int sisr_weight_prep_batched(const sisr_prep_layer* layers, int n_layers, void* s) {
  if (n_layers < 0 || (n_layers > 0 && !layers)) return fail(1, "weight_prep_batched: bad arguments");
  return wrap(weight_prep_batched(reinterpret_cast<const PrepLayer*>(layers), n_layers, S(s)),
              "weight_prep_batched");
}
int sisr_weight_prep(const float* w, const float* sigma, const float* bias, sisr_bf16* wf, sisr_bf16* wd,
                     float* bias_perm, int cout, int cin, int k, int ps_r, void* s) {
  return wrap(weight_prep(w, sigma, bias, B(wf), B(wd), bias_perm, cout, cin, k, k, ps_r, S(s)),
              "weight_prep");
}
int sisr_weight_grad_finish(const float* gp, const float* w, const float* u, const float* v,
                            const float* sigma, float* dw, const float* dbias_perm, float* dbias,
                            int cout, int cin, int k, int ps_r, int accumulate, float* ws, void* s) {
  return wrap(weight_grad_finish(gp, w, u, v, sigma, dw, dbias_perm, dbias, cout, cin, k, k, ps_r,
                                 accumulate, ws, S(s)),
              "weight_grad_finish");
}

// ------------------------------------------------------------------ convolutions
int sisr_conv_fprop(const sisr_conv_desc* d, const sisr_bf16* x, const sisr_bf16* w, const float* bias,
                    int act, float slope, const float* slope_ptr, sisr_bf16* y, float* y_nchw_f32,
                    float* stats, void* s) {
  if (!desc_ok(d)) return fail(1, "conv_fprop: inconsistent descriptor");
  if (act == SISR_ACT_PRELU && !slope_ptr) return fail(1, "conv_fprop: PReLU needs slope_ptr");
  if (tc_shape(d) && y && !y_nchw_f32 && act != SISR_ACT_TANH) {
    IgemmProblem p{};
    p.x = B(x); p.NB = d->n; p.H = d->h; p.W = d->w; p.Cin = d->cin;
    p.GH = d->oh; p.GW = d->ow; p.trav_stride = d->stride;
    p.lower_w = p.lower_h = -1;
    p.upper_w = p.upper_h = -1;
    p.w = B(w); p.Cout = d->cout; p.Ktot = 9 * d->cin; p.num_taps = 9;
    for (int t = 0; t < 9; ++t) {
      p.taps.off_h[t] = t / 3;
      p.taps.off_w[t] = t % 3;
      p.taps.k_off[t] = t * d->cin;
    }
    p.out = B(y);
    if (d->ps_r == 2) {
      p.OH = d->oh * 2; p.OW = d->ow * 2; p.ldc = d->cout / 4; p.ps_c = d->cout / 4;
    } else {
      p.OH = d->oh; p.OW = d->ow; p.ldc = d->cout; p.ps_c = 0;
    }
    p.osy = p.osx = 1; p.opy = p.opx = 0;
    p.bias = bias; p.act = act; p.slope = slope; p.slope_ptr = slope_ptr; p.stats = stats;
    p.stats_rows = igemm_max_ctas();
    if (int rc = igemm_launch(p, S(s))) return fail(rc, "conv_fprop: %s", igemm_last_error());
    return 0;
  }
  // CUDA-core paths: totals in row 0, the other rows zero
  if (stats) cudaMemsetAsync(stats, 0, sizeof(float) * 2 * d->cout * igemm_max_ctas(), S(s));
  if (d->ps_r == 2) return fail(1, "conv_fprop: PixelShuffle store needs a tensor-core shape");
  if (thin_geometry(d) && d->cin == 3 && y && !y_nchw_f32 && act != SISR_ACT_TANH) {
    const ThinConv t = thin_of(d, d->cin, d->cout);
    if (thin_in_supported(t)) {
      if (int rc = thin_in_conv(t, B(x), B(w), bias, act, slope, slope_ptr, 0, B(y), S(s)))
        return fail(rc, "conv_fprop: %s", thin_last_error());
      if (stats)
        return wrap(col_stats(B(y), static_cast<long long>(d->n) * d->oh * d->ow, d->cout, stats, 1, S(s)),
                    "col_stats");
      return 0;
    }
  }
  if (thin_geometry(d) && d->cout == 3 && !stats && act != SISR_ACT_PRELU) {
    const ThinConv t = thin_of(d, d->cout, d->cin);
    if (thin_out_supported(t)) {
      if (int rc = thin_out_conv(t, B(x), B(w), bias, act, slope, 0, B(y), y_nchw_f32, S(s)))
        return fail(rc, "conv_fprop: %s", thin_last_error());
      return 0;
    }
  }
  if (int rc = conv_fprop_simt(simt_of(d), B(x), B(w), bias, act, slope, slope_ptr, B(y), y_nchw_f32,
                               S(s)))
    return wrap(rc, "conv_fprop_simt");
  if (stats) {
    if (!y) return fail(1, "conv_fprop: stats need the bf16 output");
    return wrap(col_stats(B(y), static_cast<long long>(d->n) * d->oh * d->ow, d->cout, stats, 1, S(s)),
                "col_stats");
  }
  return 0;
}

static int conv_dgrad_impl(const sisr_conv_desc* d, const sisr_bf16* dy, const sisr_bf16* w_fprop,
                           const sisr_bf16* w_dgrad, sisr_bf16* dx, const sisr_bf16* mask, float mask_slope,
                           void* s) {
  if (!desc_ok(d)) return fail(1, "conv_dgrad: inconsistent descriptor");
  const bool even = (d->h % 2 == 0) && (d->w % 2 == 0);
  if (mask && !(tc_shape(d) && w_dgrad && (d->stride == 1 || even) && d->ps_r < 2))
    return fail(1, "conv_dgrad_masked: only on the tensor-core path (see sisr_conv_dgrad_fuses_mask)");
  if (tc_shape(d) && w_dgrad && (d->stride == 1 || even)) {
    IgemmProblem p{};
    p.mask = B(mask); p.mask_slope = mask_slope;
    p.w = B(w_dgrad); p.Cout = d->cin; p.Ktot = 9 * d->cout;
    p.out = B(dx); p.OH = d->h; p.OW = d->w; p.ldc = d->cin; p.ps_c = 0;
    p.bias = nullptr; p.act = ACT_NONE; p.stats = nullptr;
    p.x = B(dy); p.NB = d->n;
    if (d->stride == 1 && d->ps_r < 2) {
      p.H = d->oh; p.W = d->ow; p.Cin = d->cout;
      p.GH = d->h; p.GW = d->w; p.trav_stride = 1;
      p.lower_w = p.lower_h = -1; p.upper_w = p.upper_h = -1;
      p.num_taps = 9;
      for (int t = 0; t < 9; ++t) {
        p.taps.off_h[t] = 2 - t / 3;
        p.taps.off_w[t] = 2 - t % 3;
        p.taps.k_off[t] = t * d->cout;
      }
      p.osy = p.osx = 1; p.opy = p.opx = 0;
      if (int rc = igemm_launch(p, S(s))) return fail(rc, "conv_dgrad: %s", igemm_last_error());
      return 0;
    }
    if (d->stride == 1 && d->ps_r == 2) {
      // dy is [n, 2*oh, 2*ow, cout/4]; K = (tap, sub-pixel, channel)
      const int cps = d->cout / 4;
      if (cps != 64) return fail(1, "conv_dgrad: PixelShuffle dgrad needs cout/4 == 64");
      p.H = 2 * d->oh; p.W = 2 * d->ow; p.Cin = cps;
      p.GH = d->h; p.GW = d->w; p.trav_stride = 2;
      p.lower_w = p.lower_h = -2; p.upper_w = p.upper_h = -2;
      p.num_taps = 36;
      for (int t = 0; t < 9; ++t)
        for (int sub = 0; sub < 4; ++sub) {
          const int q = t * 4 + sub;
          p.taps.off_h[q] = 4 - 2 * (t / 3) + (sub >> 1);
          p.taps.off_w[q] = 4 - 2 * (t % 3) + (sub & 1);
          p.taps.k_off[q] = t * d->cout + sub * cps;
        }
      p.osy = p.osx = 1; p.opy = p.opx = 0;
      if (int rc = igemm_launch(p, S(s))) return fail(rc, "conv_dgrad(ps): %s", igemm_last_error());
      return 0;
    }
    // stride 2: four output-parity classes, each a stride-1 gather over dy
    p.H = d->oh; p.W = d->ow; p.Cin = d->cout;
    p.GH = d->h / 2; p.GW = d->w / 2; p.trav_stride = 1;
    p.lower_w = p.lower_h = 0; p.upper_w = p.upper_h = 0;
    p.osy = p.osx = 2;
    int nt = 0;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const int c = ph * 2 + pw;
        p.cls_tap_begin[c] = nt;
        for (int kh = 0; kh < 3; ++kh) {
          if ((ph + 1 - kh) % 2) continue;
          for (int kw = 0; kw < 3; ++kw) {
            if ((pw + 1 - kw) % 2) continue;
            p.taps.off_h[nt] = (ph + 1 - kh) / 2;
            p.taps.off_w[nt] = (pw + 1 - kw) / 2;
            p.taps.k_off[nt] = (kh * 3 + kw) * d->cout;
            ++nt;
          }
        }
        p.cls_tap_count[c] = nt - p.cls_tap_begin[c];
        p.cls_opy[c] = ph;
        p.cls_opx[c] = pw;
      }
    p.num_taps = nt;          // 9 taps in 4 parity classes (1 + 2 + 2 + 4), one launch
    p.n_classes = 4;
    p.opy = p.opx = 0;
    if (int rc = igemm_launch(p, S(s))) return fail(rc, "conv_dgrad(s2): %s", igemm_last_error());
    return 0;
  }
  if (d->ps_r == 2) return fail(1, "conv_dgrad: PixelShuffle layout needs a tensor-core shape");
  if (thin_geometry(d) && w_dgrad && d->cout == 3) {      // dx (wide) from a 3-channel dy
    const ThinConv t = thin_of(d, d->cout, d->cin);
    if (thin_in_supported(t)) {
      if (int rc = thin_in_conv(t, B(dy), B(w_dgrad), nullptr, ACT_NONE, 0.f, nullptr, 1, B(dx), S(s)))
        return fail(rc, "conv_dgrad: %s", thin_last_error());
      return 0;
    }
  }
  if (thin_geometry(d) && w_dgrad && d->cin == 3) {       // 3-channel dx from a wide dy
    const ThinConv t = thin_of(d, d->cin, d->cout);
    if (thin_out_supported(t)) {
      if (int rc = thin_out_conv(t, B(dy), B(w_dgrad), nullptr, ACT_NONE, 0.f, 1, B(dx), nullptr, S(s)))
        return fail(rc, "conv_dgrad: %s", thin_last_error());
      return 0;
    }
  }
  return wrap(conv_dgrad_simt(simt_of(d), B(dy), B(w_fprop), B(dx), S(s)), "conv_dgrad_simt");
}
int sisr_conv_dgrad(const sisr_conv_desc* d, const sisr_bf16* dy, const sisr_bf16* w_fprop,
                    const sisr_bf16* w_dgrad, sisr_bf16* dx, void* s) {
  return conv_dgrad_impl(d, dy, w_fprop, w_dgrad, dx, nullptr, 0.f, s);
}
int sisr_conv_dgrad_fuses_mask(const sisr_conv_desc* d) {
  const bool even = d && (d->h % 2 == 0) && (d->w % 2 == 0);
  return desc_ok(d) && tc_shape(d) && (d->stride == 1 || even) && d->ps_r < 2 ? 1 : 0;
}
int sisr_conv_dgrad_masked(const sisr_conv_desc* d, const sisr_bf16* dy, const sisr_bf16* w_fprop,
                           const sisr_bf16* w_dgrad, sisr_bf16* dx, const sisr_bf16* mask, float mask_slope,
                           void* s) {
  if (!mask) return fail(1, "conv_dgrad_masked: null mask");
  return conv_dgrad_impl(d, dy, w_fprop, w_dgrad, dx, mask, mask_slope, s);
}

size_t sisr_conv_wgrad_workspace_bytes(const sisr_conv_desc* d) {
  if (!desc_ok(d)) return 0;
  return wgrad_tc_workspace_bytes(d->n, d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->k, d->stride,
                                  d->pad, d->ps_r);
}
int sisr_conv_wgrad(const sisr_conv_desc* d, const sisr_bf16* x, const sisr_bf16* dy, float* gp,
                    float* dbias_perm, void* workspace, void* s) {
  if (!desc_ok(d)) return fail(1, "conv_wgrad: inconsistent descriptor");
  if (wgrad_tc_supported(d->n, d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->k, d->stride, d->pad,
                         d->ps_r)) {
    if (int rc = wgrad_tc_launch(B(x), B(dy), gp, dbias_perm, workspace, d->n, d->h, d->w, d->cin, d->oh,
                                 d->ow, d->cout, d->stride, d->ps_r, S(s)))
      return fail(rc, "conv_wgrad: %s", wgrad_tc_last_error());
    return 0;
  }
  if (thin_geometry(d) && d->cin == 3) {                  // thin-in conv: small = x, wide = dy
    const ThinConv t = thin_of(d, d->cin, d->cout);
    if (thin_wgrad_supported(t)) {
      if (int rc = thin_wgrad(t, B(x), B(dy), +1, gp, nullptr, dbias_perm, S(s)))
        return fail(rc, "conv_wgrad: %s", thin_last_error());
      return 0;
    }
  }
  if (thin_geometry(d) && d->cout == 3) {                 // thin-out conv: small = dy, wide = x
    const ThinConv t = thin_of(d, d->cout, d->cin);
    if (thin_wgrad_supported(t)) {
      if (int rc = thin_wgrad(t, B(dy), B(x), -1, gp, dbias_perm, nullptr, S(s)))
        return fail(rc, "conv_wgrad: %s", thin_last_error());
      return 0;
    }
  }
  return wrap(conv_wgrad_simt(simt_of(d), B(x), B(dy), gp, dbias_perm, d->ps_r == 2 ? d->cout / 4 : 0, 0,
                              S(s)),
              "conv_wgrad_simt");
}

// conv_wgrad + weight_grad_finish in one call: the split-K partials go straight into the fused
// reduce / spectral-norm / layout kernel (2 launches per layer instead of 5).
size_t sisr_conv_wgrad_fused_workspace_bytes(const sisr_conv_desc* d) {
  if (!desc_ok(d)) return 0;
  const size_t full = sizeof(float) * static_cast<size_t>(d->cout) * d->k * d->k * d->cin;
  size_t part = wgrad_tc_workspace_bytes(d->n, d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->k, d->stride,
                                         d->pad, d->ps_r);
  if (part < full) part = full;
  return part + sizeof(float) * (static_cast<size_t>(d->cout) + 8);
}
int sisr_conv_wgrad_fused(const sisr_conv_desc* d, const sisr_bf16* x, const sisr_bf16* dy,
                          const float* w_orig, const float* u, const float* v, const float* sigma,
                          float* dw, const float* dbias_in, float* dbias, int accumulate, void* workspace,
                          void* s) {
  if (!desc_ok(d)) return fail(1, "conv_wgrad_fused: inconsistent descriptor");
  if (!workspace || !dw) return fail(1, "conv_wgrad_fused: null argument");
  const size_t full = sizeof(float) * static_cast<size_t>(d->cout) * d->k * d->k * d->cin;
  size_t part = wgrad_tc_workspace_bytes(d->n, d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->k, d->stride,
                                         d->pad, d->ps_r);
  if (part < full) part = full;
  float* partials = static_cast<float*>(workspace);
  float* dbias_tmp = reinterpret_cast<float*>(static_cast<char*>(workspace) + part);
  float* dot = dbias_tmp + d->cout;
  const bool need_bias = dbias && !dbias_in;
  int splits = 1;
  if (wgrad_tc_supported(d->n, d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->k, d->stride, d->pad,
                         d->ps_r) && full % 16 == 0) {
    if (int rc = wgrad_tc_launch(B(x), B(dy), nullptr, need_bias ? dbias_tmp : nullptr, workspace, d->n,
                                 d->h, d->w, d->cin, d->oh, d->ow, d->cout, d->stride, d->ps_r, S(s),
                                 &splits))
      return fail(rc, "conv_wgrad_fused: %s", wgrad_tc_last_error());
  } else {
    if (int rc = sisr_conv_wgrad(d, x, dy, partials, need_bias ? dbias_tmp : nullptr, nullptr, s)) return rc;
  }
  const float* bsrc = dbias_in ? dbias_in : dbias_tmp;
  if (d->cin % 64 == 0 && d->k == 3)
    return wrap(weight_grad_reduce_finish(partials, splits, w_orig, u, v, sigma, dw, bsrc, dbias, d->cout,
                                          d->cin, d->k, d->k, d->ps_r, accumulate, dot, S(s)),
                "weight_grad_reduce_finish");
  return wrap(weight_grad_finish(partials, w_orig, u, v, sigma, dw, bsrc, dbias, d->cout, d->cin, d->k,
                                 d->k, d->ps_r, accumulate, dot, S(s)),
              "weight_grad_finish");
}
int sisr_debug_disable_cooperative(int off) { weight_grad_disable_cooperative(off); return 0; }
int sisr_debug_wgrad_counters(long long* device_counters) { wgrad_tc_set_debug(device_counters); return 0; }

// ------------------------------------------------------------------ BatchNorm / activations / pooling
int sisr_bn_stats(const sisr_bf16* y, long long rows, int c, float* stats, void* s) {
  cudaMemsetAsync(stats, 0, sizeof(float) * 2 * c, S(s));
  return wrap(col_stats(B(y), rows, c, stats, 1, S(s)), "bn_stats");
}
int sisr_bn_finalize(const float* stats, int stats_rows, float count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                     int training, float* scale, float* shift, float* mean, float* invstd, int c,
                     void* s) {
  return wrap(bn_finalize(stats, stats_rows, count, gamma, beta, running_mean, running_var, nbt, momentum, eps,
                          training, scale, shift, mean, invstd, c, S(s)),
              "bn_finalize");
}
int sisr_bn_apply(const sisr_bf16* y, const float* scale, const float* shift, int act, float slope,
                  const float* slope_ptr, const sisr_bf16* residual, sisr_bf16* out, long long rows,
                  int c, void* s) {
  return wrap(bn_apply(B(y), scale, shift, act, slope, slope_ptr, B(residual), B(out), rows, c, S(s)),
              "bn_apply");
}
int sisr_bn_bwd_reduce(const sisr_bf16* dout, const sisr_bf16* y, const float* mean, const float* invstd,
                       const float* scale, const float* shift, int act, float slope,
                       const float* slope_ptr, float* sums, long long rows, int c, void* s) {
  return wrap(bn_bwd_reduce(B(dout), B(y), mean, invstd, scale, shift, act, slope, slope_ptr, sums, rows,
                            c, S(s)),
              "bn_bwd_reduce");
}
int sisr_bn_bwd_apply(const sisr_bf16* dout, const sisr_bf16* y, const float* mean, const float* invstd,
                      const float* scale, const float* shift, int act, float slope,
                      const float* slope_ptr, const float* sums, float count, sisr_bf16* dy,
                      float* colsum, long long rows, int c, void* s) {
  return wrap(bn_bwd_apply(B(dout), B(y), mean, invstd, scale, shift, act, slope, slope_ptr, sums, count,
                           B(dy), colsum, rows, c, S(s)),
              "bn_bwd_apply");
}
int sisr_act_bwd(const sisr_bf16* dout, const sisr_bf16* out, int act, float slope, const float* slope_ptr,
                 sisr_bf16* din, float* dslope, float* colsum, long long rows, int c, void* s) {
  return wrap(act_bwd(B(dout), B(out), act, slope, slope_ptr, B(din), dslope, colsum, rows, c, S(s)),
              "act_bwd");
}
int sisr_maxpool2_fwd(const sisr_bf16* x, sisr_bf16* y, int n, int h, int w, int c, void* s) {
  return wrap(maxpool2_fwd(B(x), B(y), n, h, w, c, S(s)), "maxpool2_fwd");
}
int sisr_maxpool2_bwd(const sisr_bf16* x, const sisr_bf16* dy, sisr_bf16* dx, int n, int h, int w, int c,
                      void* s) {
  return wrap(maxpool2_bwd(B(x), B(dy), B(dx), n, h, w, c, S(s)), "maxpool2_bwd");
}

// ------------------------------------------------------------------ discriminator head
int sisr_dhead_forward(const sisr_bf16* x_flat, const float* w0, const float* b0, const float* w2,
                       const float* b2, float slope, float* h, float* p, int batch, int fc_in, int fc_mid,
                       void* s) {
  return wrap(dhead_forward(B(x_flat), w0, b0, w2, b2, slope, h, p, batch, fc_in, fc_mid, S(s)),
              "dhead_forward");
}
int sisr_dhead_backward(const sisr_bf16* x_flat, const float* w0, const float* w2, const float* h,
                        const float* p, const float* dp, float slope, float* dh, float* dw0, float* db0,
                        float* dw2, float* db2, float* dx_flat, int batch, int fc_in, int fc_mid,
                        int need_wgrad, void* s) {
  return wrap(dhead_backward(B(x_flat), w0, w2, h, p, dp, slope, dh, dw0, db0, dw2, db2, dx_flat, batch,
                             fc_in, fc_mid, need_wgrad, S(s)),
              "dhead_backward");
}

// ------------------------------------------------------------------ stand-alone PixelShuffle(2)
int sisr_pixel_shuffle2(const sisr_bf16* x, sisr_bf16* y, int n, int h, int w, int c_out, int inverse,
                        void* s) {
  if (!x || !y) return fail(1, "pixel_shuffle2: null argument");
  return wrap(pixel_shuffle2(B(x), B(y), n, h, w, c_out, inverse, S(s)), "pixel_shuffle2");
}

// ------------------------------------------------------------------ LR synthesis
int sisr_lr_from_hr(const float* hr, float* lr, int n, int c, int h, int w, int oh, int ow, void* s) {
  if (!hr || !lr) return fail(1, "lr_from_hr: null argument");
  return wrap(lr_from_hr(hr, lr, nullptr, nullptr, n, c, h, w, oh, ow, S(s)), "lr_from_hr");
}
int sisr_lr_from_hr_bwd(const float* hr, const float* dlr, float* dhr, int n, int c, int h, int w, int oh,
                        int ow, void* s) {
  if (!hr || !dlr || !dhr) return fail(1, "lr_from_hr_bwd: null argument");
  return wrap(lr_from_hr(hr, nullptr, dlr, dhr, n, c, h, w, oh, ow, S(s)), "lr_from_hr_bwd");
}

// ------------------------------------------------------------------ losses
int sisr_bce_fwd(const float* p, int n, float target, float* loss, float* mean_p, void* s) {
  return wrap(bce_fwd(p, n, target, loss, mean_p, S(s)), "bce_fwd");
}
int sisr_bce_bwd(const float* p, int n, float target, const float* gout, float* dp, void* s) {
  return wrap(bce_bwd(p, n, target, gout, dp, S(s)), "bce_bwd");
}
int sisr_mse_fwd(const float* a, const float* b, long long n, float coef, float* loss, void* s) {
  return wrap(mse_fwd(a, b, n, coef, loss, S(s)), "mse_fwd");
}
int sisr_mse_bwd(const float* a, const float* b, long long n, float coef, const float* gout, float* ga,
                 float* gb, void* s) {
  return wrap(mse_bwd(a, b, n, coef, gout, ga, gb, S(s)), "mse_bwd");
}

// ------------------------------------------------------------------ image-quality metrics
int sisr_psnr_ssim(const float* a, const float* b, int n, int c, int h, int w, float range, float* workspace,
                   float* psnr, float* ssim, void* s) {
  if (!a || !b || !workspace || !psnr || !ssim) return fail(1, "psnr_ssim: null argument");
  if (h < 11 || w < 11) return fail(1, "psnr_ssim: images must be at least 11 x 11 (SSIM window)");
  return wrap(psnr_ssim(a, b, n, c, h, w, range, workspace, psnr, ssim, S(s)), "psnr_ssim");
}

// ------------------------------------------------------------------ SyncBN over NVLink peer memory
size_t sisr_peer_workspace_bytes(void) { return peer_workspace_bytes(); }
int sisr_peer_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }
int sisr_peer_alloc(void** ptr) {
  if (!ptr) return fail(1, "peer_alloc: null argument");
  cudaError_t e = cudaMalloc(ptr, peer_workspace_bytes());
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, peer_workspace_bytes());
  if (e != cudaSuccess) return fail(2, "peer_alloc: %s", cudaGetErrorString(e));
  return 0;
}
int sisr_peer_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? 0 : fail(2, "peer_free failed"); }
int sisr_peer_get_handle(void* ptr, void* handle) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return fail(2, "peer_get_handle: %s", cudaGetErrorString(e));
  memcpy(handle, &h, sizeof h);
  return 0;
}
int sisr_peer_open(const void* handle, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(2, "peer_open: %s", cudaGetErrorString(e));
  return 0;
}
int sisr_peer_close(void* ptr) {
  return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? 0 : fail(2, "peer_close failed");
}
static PeerTable make_table(void* const* bases, int rank, int world) {
  PeerTable t{};
  t.rank = rank;
  t.world = world;
  for (int r = 0; r < world && r < kPeerMaxWorld; ++r) t.base[r] = bases[r];
  return t;
}
int sisr_peer_allreduce(void* const* bases, int rank, int world, int slot, float* buf, int n, void* s) {
  if (!bases || world > kPeerMaxWorld) return fail(1, "peer_allreduce: bad arguments");
  return wrap(peer_allreduce(make_table(bases, rank, world), slot, buf, n, S(s)), "peer_allreduce");
}
int sisr_bn_finalize_sync(void* const* bases, int rank, int world, int slot, const float* stats,
                          int stats_rows, float count, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* nbt, float momentum,
                          float eps, float* scale, float* shift, float* mean, float* invstd, int c,
                          void* s) {
  if (!bases || world > kPeerMaxWorld) return fail(1, "bn_finalize_sync: bad arguments");
  return wrap(bn_finalize_sync(make_table(bases, rank, world), slot, stats, stats_rows, count, gamma,
                               beta, running_mean, running_var, nbt, momentum, eps, scale, shift, mean,
                               invstd, c, S(s)),
              "bn_finalize_sync");
}

/* sisr_bn_bwd_reduce + the cross-GPU exchange of its [2c+1] sums in one launch: the last CTA to finish
 * runs the NVLink peer exchange and writes the global sums to sums_global. */
int sisr_bn_bwd_reduce_sync(void* const* bases, int rank, int world, int slot, const sisr_bf16* dout,
                            const sisr_bf16* y, const float* mean, const float* invstd, const float* scale,
                            const float* shift, int act, float slope, const float* slope_ptr, float* sums,
                            float* sums_global, void* ticket, long long rows, int c, void* s) {
  if (!bases || world > kPeerMaxWorld || !sums_global || !ticket)
    return fail(1, "bn_bwd_reduce_sync: bad arguments");
  const PeerTable t = make_table(bases, rank, world);
  return wrap(bn_bwd_reduce(B(dout), B(y), mean, invstd, scale, shift, act, slope, slope_ptr, sums, rows, c,
                            S(s), &t, slot, sums_global, static_cast<unsigned int*>(ticket)),
              "bn_bwd_reduce_sync");
}

// ------------------------------------------------------------------ optimiser
int sisr_adam_tick(int* step, float lr0, float decay, float b1, float b2, float* hyper, void* s) {
  return wrap(adam_tick(step, lr0, decay, b1, b2, hyper, S(s)), "adam_tick");
}
int sisr_adam_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                    const long long* numel, const float* hyper, float b1, float b2, float eps,
                    float grad_scale, void* s) {
  return wrap(adam_multi(n, p, g, m, v, numel, hyper, b1, b2, eps, grad_scale, S(s)), "adam_multi");
}

}  // extern "C"
