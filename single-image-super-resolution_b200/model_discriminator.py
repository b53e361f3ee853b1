"""Drop-in ``Discriminator`` (reference: model_discriminator.py:5-76).

conv3x3(3->64)+LeakyReLU, 7 x [spectral-norm conv3x3 (stride 1/2) -> BatchNorm -> LeakyReLU],
flatten in the reference's (c,h,w) order, Linear -> LeakyReLU -> Linear -> Sigmoid; output (B, 1).
Same constructor, asserts, attributes (``fc_in``, ``fc_mid``, ``conv``, ``fc``), lenient
``load_state_dict`` and ``state_dict`` keys as the reference.
"""
from __future__ import annotations

import torch.nn as nn

from . import ops
from .layers import ACT_LEAKY, SNConv2d, bn_act


class BasicBlock(nn.Module):
    """Non-residual block of D (model_discriminator.py:5-15)."""

    def __init__(self, n_in, n_out, stride):
        super().__init__()
        self.layers = nn.Sequential(
            SNConv2d(n_in, n_out, 3, stride, 1, sn=True),
            nn.BatchNorm2d(n_out),
            nn.LeakyReLU())

    def forward_nhwc(self, x):
        conv, bn, _ = self.layers
        y, st = conv.run(x, want_stats=bn.training)
        return bn_act(bn, y, st, act=ACT_LEAKY)

    def forward(self, x):
        return ops.ToNCHW.apply(self.forward_nhwc(ops.ToNHWC.apply(x)))


class Discriminator(nn.Module):
    def __init__(self, input_shape, list_n_features, list_stride):
        super().__init__()
        channels, width, height = input_shape
        widths, strides = list(list_n_features), list(list_stride)
        # the reference's three constructor checks (model_discriminator.py:27-32), same exception type
        if any(s not in (1, 2) for s in strides):
            raise AssertionError("strides must be 1 or 2 (SRGAN uses no other)")
        shrink = 4 ** sum(1 for s in strides if s == 2)        # every stride-2 conv quarters the pixel count
        if (width * height) % shrink:
            raise AssertionError("input size is not divisible by the total down-sampling factor")
        if len(widths) != len(strides):
            raise AssertionError("list_n_features and list_stride differ in length")
        self.fc_in = width * height * widths[-1] // shrink     # flattened size entering the head
        self.fc_mid = 2 * widths[-1]
        stem = SNConv2d(channels, widths[0], 3, strides[0], 1, sn=True)
        stages = [BasicBlock(n_in, n_out, st) for n_in, n_out, st in zip(widths[:-1], widths[1:], strides[1:])]
        self.conv = nn.Sequential(stem, nn.LeakyReLU(), nn.Sequential(*stages))
        self.fc = nn.Sequential(nn.Linear(self.fc_in, self.fc_mid), nn.LeakyReLU(),
                                nn.Linear(self.fc_mid, 1), nn.Sigmoid())

    def forward(self, x):
        ops.prepare_convs([self.conv[0]] + [b.layers[0] for b in self.conv[2]], x.requires_grad)
        x = ops.ToNHWC.apply(x)
        x, _ = self.conv[0].run(x, act=ACT_LEAKY)
        for block in self.conv[2]:
            x = block.forward_nhwc(x)
        if x.shape[1] * x.shape[2] * x.shape[3] != self.fc_in:
            raise RuntimeError(f"discriminator input does not match fc_in={self.fc_in}")
        return ops.DHeadFn.apply(x, self.fc[0].weight, self.fc[0].bias, self.fc[2].weight,
                                 self.fc[2].bias)

    def load_state_dict(self, state_dict, strict=True):
        """strict: ``nn.Module`` behaviour.  Otherwise (model_discriminator.py:64-76) every entry whose name
        exists here is copied in place; unknown names are skipped and a failing copy (shape mismatch) is
        reported on stdout instead of raised."""
        if strict:
            return nn.Module.load_state_dict(self, state_dict, strict)
        mine = self.state_dict()
        for key in (k for k in state_dict if k in mine):
            try:
                mine[key].copy_(state_dict[key])
            except Exception as err:  # noqa: BLE001 - lenient by contract
                print(f"dis: could not load {key}: {err}")
