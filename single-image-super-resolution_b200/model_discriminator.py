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
        w, h = input_shape[1], input_shape[2]
        for s in list_stride:
            assert s in (1, 2), "l'article utilise des stride de 1 ou 2 seulement"
        assert w * h % 4 ** (sum(list_stride) - len(list_stride)) == 0, \
            "chaque stride à 2 divisise la taille par 2, il faut que ca soit divisible"
        assert len(list_n_features) == len(list_stride)
        self.fc_in = w * h * list_n_features[-1] // (4 ** (sum(list_stride) - len(list_stride)))
        self.fc_mid = list_n_features[-1] * 2
        self.conv = nn.Sequential(
            SNConv2d(input_shape[0], list_n_features[0], 3, list_stride[0], 1, sn=True),
            nn.LeakyReLU(),
            nn.Sequential(*[BasicBlock(list_n_features[i - 1], list_n_features[i], list_stride[i])
                            for i in range(1, len(list_n_features))]))
        self.fc = nn.Sequential(
            nn.Linear(self.fc_in, self.fc_mid),
            nn.LeakyReLU(),
            nn.Linear(self.fc_mid, 1),
            nn.Sigmoid())

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
        """strict -> default behaviour; otherwise copy every tensor whose name matches and report
        shape errors (model_discriminator.py:64-76)."""
        if strict:
            return nn.Module.load_state_dict(self, state_dict, strict)
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name not in own_state:
                continue
            try:
                own_state[name].copy_(param)
            except Exception as e:  # noqa: BLE001 - mirrors the reference's lenient loader
                print("dis: lecture échouée pour", name, "  -  ", e)
