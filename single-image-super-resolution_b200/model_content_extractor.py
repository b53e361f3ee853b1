"""Drop-in ``MaskedVGG`` / ``identity`` (reference: model_content_extractor.py:6-73).

Frozen VGG19 ``features[:k]``; the outputs tapped at the conv before each selected max-pool are
flattened (NCHW order) and concatenated, shape (B, sum of taps).  The reference's in-place-ReLU
behaviour is kept: every tap but the last is seen after the following ReLU, the last tap is the
raw conv output.  conv+bias+ReLU is one tcgen05 kernel per layer.

Pretrained ImageNet weights cannot be downloaded in this environment; the layer stack is created
with torchvision's default initialisation and ``load_state_dict`` accepts the usual
``layers.{idx}.weight/bias`` keys (e.g. sliced from a local vgg19 checkpoint).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .layers import ACT_NONE, ACT_RELU
from .ops import Conv2dFn, ConvCfg, MaxPool2Fn

# indices of the max-pool layers in torchvision's VGG19 ``features`` (the last one is never used)
maxPool_indexes = (4, 9, 18, 27, 36)
maxPool_indexes_before_act = [x - 1 for x in maxPool_indexes]
layersSize = (64, 128, 256, 512, 512)

_VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
              512, 512, 512, 512, "M"]


def identity():
    """plain pixel MSE between the two images"""
    return nn.Identity()


def _vgg19_features():
    layers, cin = [], 3
    for v in _VGG19_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            conv = nn.Conv2d(cin, v, kernel_size=3, padding=1)
            nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
            nn.init.constant_(conv.bias, 0)
            layers += [conv, nn.ReLU(inplace=True)]
            cin = v
    return nn.Sequential(*layers)


class MaskedVGG(nn.Module):
    def __init__(self, mask):
        super().__init__()
        self.intermediate_layers_kept = [maxPool_indexes_before_act[i]
                                         for i in range(len(maxPool_indexes_before_act)) if mask & (1 << i)]
        self.layers = _vgg19_features()[:self.intermediate_layers_kept[-1]]
        self.layers.eval()
        self.layers.requires_grad = False
        for param in self.layers.parameters():
            param.requires_grad = False
        self._prepared = {}

    def _prep(self, idx, conv, need_dgrad):
        """bf16 [Cout,3,3,Cin] / [Cin,3,3,Cout] copies of the frozen weights, rebuilt only when the
        parameter changes (load_state_dict, .to())."""
        w = conv.weight
        key = (w.data_ptr(), w._version, w.device)
        hit = self._prepared.get(idx)
        if hit is None or hit[0] != key or (need_dgrad and hit[2] is None):
            cout, cin, k, _ = w.shape
            wf = torch.empty((cout, k, k, cin), dtype=torch.bfloat16, device=w.device)
            wd = torch.empty((cin, k, k, cout), dtype=torch.bfloat16, device=w.device)
            ops.call("sisr_weight_prep", w, None, None, wf, wd, None, cout, cin, k, 0, ops._stream())
            hit = (key, wf, wd)
            self._prepared[idx] = hit
        return hit[1], hit[2]

    def forward(self, x):
        x = ops.ToNHWC.apply(x)
        kept = self.intermediate_layers_kept
        saved = []
        n = len(self.layers)
        for i, layer in enumerate(self.layers, 1):
            if isinstance(layer, nn.Conv2d):
                fuse_relu = i < n  # a ReLU follows unless this is the last (tapped) conv
                cfg = ConvCfg(stride=1, pad=1, act=ACT_RELU if fuse_relu else ACT_NONE, training=False)
                wf, wd = self._prep(i, layer, x.requires_grad)
                x, _ = Conv2dFn.apply(x, layer.weight, layer.bias, None, None, None, cfg, (wf, wd))
                if i in kept:
                    saved.append(x)
            elif isinstance(layer, nn.MaxPool2d):
                x = MaxPool2Fn.apply(x)
            # nn.ReLU: already applied in the conv epilogue
        return torch.cat([ops.ToNCHW.apply(e).view(e.shape[0], -1) for e in saved], dim=1)


def get_size(im, mask):
    assert im.shape[1] == 3
    w, h = im.shape[2], im.shape[3]
    size = 0
    for i in range(len(layersSize)):
        if mask & (1 << i):
            size += (w // 2 ** i) * (h // 2 ** i) * layersSize[i]
    return size
