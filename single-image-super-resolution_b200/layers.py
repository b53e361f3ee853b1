"""Parameter-holding building blocks shared by the drop-in modules.

The reference builds its networks from ``nn.Sequential`` containers of ``nn.Conv2d`` (wrapped by the
legacy ``torch.nn.utils.spectral_norm`` hook), ``nn.BatchNorm2d`` and ``nn.PReLU``; its
``state_dict`` key layout is part of the drop-in contract.  Here the same
containers hold the parameters, but the arithmetic is driven by the parent module, which chains
the fused CUDA operators of :mod:`sisr_b200.ops` over NHWC bf16 activations.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH, BnActFn, BnCfg, Conv2dFn, ConvCfg


class SNConv2d(nn.Module):
    """Conv2d parameters, optionally under the legacy spectral-norm parameterisation.

    With ``sn=True`` the registered names are exactly those the legacy hook produces on an
    ``nn.Conv2d``: parameters ``bias``, ``weight_orig`` and buffers ``weight_u`` ([Cout]) and
    ``weight_v`` ([Cin*k*k]); without it, ``weight`` and ``bias``.
    """

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1,
                 padding: int = 0, sn: bool = False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.sn = kernel_size, stride, padding, sn
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))          # nn.Conv2d.reset_parameters
        bound = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
        b = torch.empty(out_channels).uniform_(-bound, bound)
        if sn:
            self.bias = nn.Parameter(b)
            self.weight_orig = nn.Parameter(w)
            u = torch.randn(out_channels)
            v = torch.randn(in_channels * kernel_size * kernel_size)
            self.register_buffer("weight_u", u / u.norm().clamp_min(ops.SN_EPS))
            self.register_buffer("weight_v", v / v.norm().clamp_min(ops.SN_EPS))
        else:
            self.weight = nn.Parameter(w)
            self.bias = nn.Parameter(b)
        self._prep = None
        self.ps_r = 0          # 2 when a PixelShuffle(2) follows (set by the owning stage)

    @property
    def master_weight(self) -> torch.Tensor:
        return self.weight_orig if self.sn else self.weight

    def run(self, x: torch.Tensor, *, act: int = ACT_NONE, slope: Optional[torch.Tensor] = None,
            ps_r: int = 0, want_stats: bool = False, out_nchw_f32: bool = False):
        """x: NHWC bf16.  Returns (y, stats)."""
        cfg = ConvCfg(stride=self.stride, pad=self.padding, act=act, ps_r=ps_r,
                      want_stats=want_stats, training=self.training, out_nchw_f32=out_nchw_f32)
        u = self.weight_u if self.sn else None
        v = self.weight_v if self.sn else None
        prep, self._prep = self._prep, None      # set by ops.prepare_convs for exactly one forward
        return Conv2dFn.apply(x, self.master_weight, self.bias, u, v, slope, cfg, prep)

    def forward(self, x):  # module-boundary use: NCHW fp32 in / out
        y, _ = self.run(ops.ToNHWC.apply(x))
        return ops.ToNCHW.apply(y)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, padding={self.padding}, spectral_norm={self.sn}")


def bn_act(bn: nn.BatchNorm2d, y: torch.Tensor, stats, *, act: int = ACT_NONE,
           slope: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None):
    """Apply the BatchNorm held by ``bn`` (+activation, +residual) to NHWC bf16 ``y``."""
    cfg = BnCfg(act=act, training=bn.training, momentum=bn.momentum, eps=bn.eps)
    return BnActFn.apply(y, stats, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                         bn.num_batches_tracked, residual, slope, cfg)


__all__ = ["SNConv2d", "bn_act", "ACT_NONE", "ACT_RELU", "ACT_LEAKY", "ACT_PRELU", "ACT_TANH"]
