"""Drop-in ``Generator`` / ``GeneratorSuffix`` (reference: model_generator.py:5-141).

Same constructors, attributes (``first_layers``, ``block_list``, ``block_list_end``, ``upscale``,
``end``, ``base``, ``n_features_last``), methods (``forward``, ``forward_no_end``, ``freeze``,
``load_state_dict`` with the coverage report) and ``state_dict`` keys as the reference, so that
``config.py`` / ``train.py`` / ``visualisation.py`` and reference checkpoints work unchanged.
Underneath, the forward pass is a chain of fused sm_100a kernels over NHWC bf16 activations:

  conv9x9 + PReLU | 16 x [conv3x3 (+BN stats) -> BN+PReLU -> conv3x3 (+BN stats) -> BN + skip]
  | conv3x3 -> BN + global skip | conv3x3 64->256 + PReLU with PixelShuffle folded into the store
  | conv3x3 -> Tanh written as NCHW fp32 at the module boundary.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .layers import ACT_NONE, ACT_PRELU, ACT_TANH, SNConv2d, bn_act


class BasicBlock(nn.Module):
    """Residual block of G (model_generator.py:5-19): x + BN(conv(PReLU(BN(conv(x)))))."""

    def __init__(self, n_features):
        super().__init__()
        self.layers = nn.Sequential(
            SNConv2d(n_features, n_features, 3, 1, 1, sn=True),
            nn.BatchNorm2d(n_features),
            nn.PReLU(),
            SNConv2d(n_features, n_features, 3, 1, 1, sn=True),
            nn.BatchNorm2d(n_features))

    def forward_nhwc(self, x):
        c1, bn1, act, c2, bn2 = self.layers
        y, st = c1.run(x, want_stats=bn1.training)
        a = bn_act(bn1, y, st, act=ACT_PRELU, slope=act.weight)
        y, st = c2.run(a, want_stats=bn2.training)
        return bn_act(bn2, y, st, residual=x)

    def forward(self, x):
        return ops.ToNCHW.apply(self.forward_nhwc(ops.ToNHWC.apply(x)))


class _UpscaleStage(nn.Sequential):
    """[conv3x3 -> PixelShuffle -> PReLU]; the three are one kernel (shuffle = store addressing,
    the single-slope PReLU commutes with it)."""

    def __init__(self, cin, cout, scale, sn):
        if scale != 2:
            raise NotImplementedError("only PixelShuffle(2) stages are built (the reference's "
                                      "config.py uses list_scales = [2])")
        super().__init__(SNConv2d(cin, cout, 3, 1, 1, sn=sn), nn.PixelShuffle(scale), nn.PReLU())
        self[0].ps_r = scale

    def forward_nhwc(self, x):
        y, _ = self[0].run(x, act=ACT_PRELU, slope=self[2].weight, ps_r=2)
        return y

    def forward(self, x):
        return ops.ToNCHW.apply(self.forward_nhwc(ops.ToNHWC.apply(x)))


class _EndStage(nn.Sequential):
    """[conv3x3 -> Tanh] producing the NCHW fp32 image."""

    def __init__(self, cin, cout, sn):
        super().__init__(SNConv2d(cin, cout, 3, 1, 1, sn=sn), nn.Tanh())

    def forward_nhwc(self, x):
        y, _ = self[0].run(x, act=ACT_TANH, out_nchw_f32=True)
        return y

    def forward(self, x):
        return self.forward_nhwc(ops.ToNHWC.apply(x))


class Generator(nn.Module):
    def __init__(self, n_blocks, n_features_block, n_features_last, list_scales, use_sn=False,
                 input_channels=3):
        super().__init__()
        assert n_features_last % 4 == 0
        self.n_features_last = n_features_last
        self.first_layers = nn.Sequential(
            SNConv2d(input_channels, n_features_block, 9, 1, 4, sn=True), nn.PReLU())
        self.block_list = nn.Sequential(*[BasicBlock(n_features_block) for _ in range(n_blocks)])
        self.block_list_end = nn.Sequential(
            SNConv2d(n_features_block, n_features_block, 3, 1, 1, sn=True),
            nn.BatchNorm2d(n_features_block))
        stages = []
        for i, s in enumerate(list_scales):
            cin = n_features_block if i == 0 else n_features_last // list_scales[i - 1] ** 2
            stages.append(_UpscaleStage(cin, n_features_last, s, use_sn))
        self.upscale = nn.Sequential(*stages)
        self.end = _EndStage(n_features_last // list_scales[-1] ** 2, input_channels, use_sn)

    # -- reference API ---------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=False):
        """Non-strict load plus the reference's coverage report (model_generator.py:65-84)."""
        result = super().load_state_dict(state_dict, strict=strict)
        mine, theirs = self.state_dict(), state_dict
        same_keys = mine.keys() == theirs.keys()
        if not same_keys or any(torch.any(mine[k].cpu() != theirs[k].cpu()) for k in mine):
            n_mine = sum(t.nelement() for t in mine.values())
            n_theirs = sum(t.nelement() for t in theirs.values())
            n_common = sum(mine[k].nelement() for k in set(mine) & set(theirs))
            print("chargement du générateur à ", round(n_common / n_mine * 100, 1), "%",
                  "    (", round(n_common * 1e-6, 2), " M)", sep="")
            print("  - architecture : ", len(mine), " ens de poids (", round(n_mine * 1e-6, 2), " M)", sep="")
            print("  - checkpoint   : ", len(theirs), " ens de poids (", round(n_theirs * 1e-6, 2), " M)", sep="")
            missing = mine.keys() - theirs.keys()
            print("  - manquants    :", len(missing), missing)
            unused = theirs.keys() - mine.keys()
            print("  - non utilisés :", len(unused), unused)
        return result

    def convs_no_end(self):
        """The SNConv2d modules in forward order (without the output conv)."""
        convs = [self.first_layers[0]]
        for block in self.block_list:
            convs += [block.layers[0], block.layers[3]]
        convs.append(self.block_list_end[0])
        convs += [stage[0] for stage in self.upscale]
        return convs

    def forward_no_end_nhwc(self, x, tap_prefix=""):
        """x: NHWC bf16 LR image -> NHWC bf16 feature map after the upscale stages."""
        conv0, act0 = self.first_layers
        x, _ = conv0.run(x, act=ACT_PRELU, slope=act0.weight)
        skip = ops.tap(tap_prefix + "first_layers", x)
        for i, block in enumerate(self.block_list):
            x = ops.tap(f"{tap_prefix}block_list.{i}", block.forward_nhwc(x))
        conv_e, bn_e = self.block_list_end
        y, st = conv_e.run(x, want_stats=bn_e.training)
        x = ops.tap(tap_prefix + "block_list_end", bn_act(bn_e, y, st, residual=skip))
        for s, stage in enumerate(self.upscale):
            x = ops.tap(f"{tap_prefix}upscale.{s}", stage.forward_nhwc(x))
        return x

    def forward_no_end(self, x):
        ops.prepare_convs(self.convs_no_end(), x.requires_grad)
        return ops.ToNCHW.apply(self.forward_no_end_nhwc(ops.ToNHWC.apply(x)))

    def forward(self, x):
        ops.prepare_convs(self.convs_no_end() + [self.end[0]], x.requires_grad)
        return self.end.forward_nhwc(self.forward_no_end_nhwc(ops.ToNHWC.apply(x)))

    def freeze(self, freeze_upscale=False, freeze_end=False):
        layer_list = [self.first_layers, self.block_list, self.block_list_end]
        if freeze_upscale:
            layer_list.append(self.upscale)
        if freeze_end:
            layer_list.append(self.end)
        for layer in layer_list:
            layer.requires_grad = False
            for p in layer.parameters():
                p.requires_grad = False


class GeneratorSuffix(nn.Module):
    """One more x2 stage on top of a trained generator (model_generator.py:117-141); the output
    conv of the wrapped generator is reused through a plain python list so that it is registered
    only once (under ``base.``)."""

    def __init__(self, prefix, freeze_prefix=False, **kwargs):
        super().__init__()
        self.base = prefix
        self.n_features_last = prefix.n_features_last
        self.upscale = _UpscaleStage(self.n_features_last // 4, self.n_features_last, 2, True)
        self.end = [prefix.end[0] if type(prefix.end) == list else prefix.end]
        if freeze_prefix:
            prefix.freeze(**kwargs)

    def freeze(self, freeze_upscale=False, freeze_end=False):
        # lets a suffix be wrapped (and frozen) by a further suffix; the reference only defines
        # freeze() on Generator, so this follows the same contract one level up
        self.base.freeze(freeze_upscale=True, freeze_end=freeze_end)
        if freeze_upscale:
            for p in self.upscale.parameters():
                p.requires_grad = False

    def convs_no_end(self):
        return self.base.convs_no_end() + [self.upscale[0]]

    def forward_no_end_nhwc(self, x, tap_prefix=""):
        x = self.base.forward_no_end_nhwc(x, tap_prefix + "base.")
        return ops.tap(tap_prefix + "upscale", self.upscale.forward_nhwc(x))

    def forward_no_end(self, x):
        ops.prepare_convs(self.convs_no_end(), x.requires_grad)
        return ops.ToNCHW.apply(self.forward_no_end_nhwc(ops.ToNHWC.apply(x)))

    def forward(self, x):
        ops.prepare_convs(self.convs_no_end() + [self.end[0][0]], x.requires_grad)
        return self.end[0].forward_nhwc(self.forward_no_end_nhwc(ops.ToNHWC.apply(x)))
