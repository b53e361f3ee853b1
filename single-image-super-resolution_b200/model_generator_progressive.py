"""Drop-in for the reference's older progressive generator (model_generator_progressive.py:4-65;
not imported by config.py / train.py / visualisation.py, kept by the reference as an alternative design).

``GeneratorProgresiveBase`` is the residual trunk alone (no spectral norm, no global skip);
``GeneratorSuffix(prefix, n_features)`` appends conv3x3(nf->nf) + PixelShuffle(2) + PReLU and a
conv3x3(nf/4 -> 3) + Tanh head, and stages are chained through ``previous.beginning`` with the channel
count shrinking by 4 per stage (64 -> 16 -> 4).  Same constructors, attribute names and ``state_dict``
keys as the reference.  The 64-channel convs run on the tcgen05 engine, the narrow ones (16, 4
channels) on the CUDA-core kernels; the shuffle of these stages is a stand-alone kernel
(``sisr_pixel_shuffle2``) because a 16-channel row is too short for the fused store.
"""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import ACT_PRELU, ACT_TANH, SNConv2d, bn_act
from .ops import Conv2dFn, ConvCfg


def _run_padded(conv: SNConv2d, x, cin_pad: int, cout_pad: int, act: int = 0, slope=None,
                out_nchw_f32: bool = False):
    """A conv whose channel counts are below the 8-channel vectors of the kernels (the 4-channel stage of
    the reference's own smoke test, model_generator_progressive.py:67-86): the parameters keep the
    reference's shapes, the kernels see zero-padded copies (``x`` already has ``cin_pad`` channels whose
    tail is zero), and autograd carries the gradients of the padding back to the real parameters."""
    cout, cin = conv.out_channels, conv.in_channels
    w = F.pad(conv.master_weight, (0, 0, 0, 0, 0, cin_pad - cin, 0, cout_pad - cout))
    b = F.pad(conv.bias, (0, cout_pad - cout))
    cfg = ConvCfg(stride=conv.stride, pad=conv.padding, act=act, training=conv.training,
                  out_nchw_f32=out_nchw_f32, sync_wgrad=True)
    y, _ = Conv2dFn.apply(x, w, b, None, None, slope, cfg, None)
    return y


class BasicBlock(nn.Module):
    """model_generator_progressive.py:4-18."""

    def __init__(self, n_features):
        super().__init__()
        self.layers = nn.Sequential(
            SNConv2d(n_features, n_features, 3, 1, 1, sn=False),
            nn.BatchNorm2d(n_features),
            nn.PReLU(),
            SNConv2d(n_features, n_features, 3, 1, 1, sn=False),
            nn.BatchNorm2d(n_features))

    def forward_nhwc(self, x):
        c1, bn1, act, c2, bn2 = self.layers
        y, st = c1.run(x, want_stats=bn1.training)
        a = bn_act(bn1, y, st, act=ACT_PRELU, slope=act.weight)
        y, st = c2.run(a, want_stats=bn2.training)
        return bn_act(bn2, y, st, residual=x)

    def forward(self, x):
        return ops.ToNCHW.apply(self.forward_nhwc(ops.ToNHWC.apply(x)))


class GeneratorProgresiveBase(nn.Module):
    """model_generator_progressive.py:21-44 (the reference's spelling is kept)."""

    def __init__(self, n_blocks, n_features, input_channels=3):
        super().__init__()
        self.first_layers = nn.Sequential(
            SNConv2d(input_channels, n_features, 9, 1, 4, sn=False), nn.PReLU())
        self.block_list = nn.Sequential(*[BasicBlock(n_features) for _ in range(n_blocks)])
        self.block_list_end = nn.Sequential(
            SNConv2d(n_features, n_features, 3, 1, 1, sn=False), nn.BatchNorm2d(n_features))

    def convs(self):
        out = [self.first_layers[0]]
        for block in self.block_list:
            out += [block.layers[0], block.layers[3]]
        return out + [self.block_list_end[0]]

    def forward_nhwc(self, x):
        conv0, act0 = self.first_layers
        x, _ = conv0.run(x, act=ACT_PRELU, slope=act0.weight)
        for block in self.block_list:
            x = block.forward_nhwc(x)
        conv_e, bn_e = self.block_list_end
        y, st = conv_e.run(x, want_stats=bn_e.training)
        return bn_act(bn_e, y, st)

    def forward(self, x):
        ops.prepare_convs(self.convs(), x.requires_grad)
        return ops.ToNCHW.apply(self.forward_nhwc(ops.ToNHWC.apply(x)))


class _Beginning(nn.Sequential):
    """[prefix, conv3x3 nf->nf, PixelShuffle(2), PReLU] (model_generator_progressive.py:52-56); the
    single-slope PReLU commutes with the shuffle and is fused into the conv epilogue."""

    @property
    def narrow(self):
        """n_features below the kernels' 8-channel vectors (the 4-channel stage)."""
        return self[1].out_channels % 8 != 0

    def convs(self):
        """convs prepared in one batch (the narrow stage prepares its zero-padded copies itself)"""
        return self[0].convs() + ([] if self.narrow else [self[1]])

    def forward_nhwc(self, x):
        x = self[0].forward_nhwc(x)
        if self.narrow:
            # 4 -> 4 channels run as 8 -> 32: after the shuffle channel 0 is the real one, 1..7 are zero
            x = F.pad(x, (0, 8 - x.shape[-1]))
            y = _run_padded(self[1], x, 8, 32, act=ACT_PRELU, slope=self[3].weight)
        else:
            y, _ = self[1].run(x, act=ACT_PRELU, slope=self[3].weight)
        return ops.PixelShuffle2Fn.apply(y)

    def forward(self, x):
        ops.prepare_convs(self.convs(), x.requires_grad)
        y = ops.ToNCHW.apply(self.forward_nhwc(ops.ToNHWC.apply(x)))
        return y[:, :self[1].out_channels // 4] if self.narrow else y


class GeneratorSuffix(nn.Module):
    def __init__(self, prefix, n_features, input_channels=3):
        super().__init__()
        assert n_features % 4 == 0
        if n_features % 8 and n_features != 4:
            raise NotImplementedError("only the 4-channel narrow stage of the reference's own test "
                                      "(model_generator_progressive.py:67-86) is built below 8 channels")
        self.beginning = _Beginning(
            prefix,
            SNConv2d(n_features, n_features, 3, 1, 1, sn=False),
            nn.PixelShuffle(upscale_factor=2),
            nn.PReLU())
        self.end = nn.Sequential(
            SNConv2d(n_features // 4, input_channels, 3, 1, 1, sn=False), nn.Tanh())

    def forward(self, x):
        narrow = self.beginning.narrow
        ops.prepare_convs(self.beginning.convs() + ([] if narrow else [self.end[0]]), x.requires_grad)
        x = self.beginning.forward_nhwc(ops.ToNHWC.apply(x))
        if narrow:          # end conv 1 -> 3 runs as 8 -> 3 on the zero-padded map
            return _run_padded(self.end[0], x, 8, 3, act=ACT_TANH, out_nchw_f32=True)
        y, _ = self.end[0].run(x, act=ACT_TANH, out_nchw_f32=True)
        return y
