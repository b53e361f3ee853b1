"""Seeded state-dict factories for the oracle and the tests.  TEST INFRASTRUCTURE ONLY.

Builds tensors under the reference's state_dict key layout from a seed with
a self-contained recipe (CPU ``torch.Generator``), so that fixtures only need to store the seed and
the reference's outputs, not the weights.  Loading these dicts into the real reference modules
with ``strict=True`` (oracle/validate_against_reference.py) is what pins the key layout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch

from .srgan_oracle import vgg19_feature_layers, vgg_kept_positions

State = Dict[str, torch.Tensor]


class _Rng:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def uniform(self, shape, bound):
        return (torch.rand(shape, generator=self.g) * 2 - 1) * bound

    def normal(self, shape, std=1.0):
        return torch.randn(shape, generator=self.g) * std


def _conv(st: State, p: str, rng: _Rng, cin: int, cout: int, k: int, sn: bool):
    bound = 1.0 / math.sqrt(cin * k * k)
    st[p + "bias"] = rng.uniform((cout,), bound)
    w = rng.uniform((cout, cin, k, k), bound)
    if sn:
        st[p + "weight_orig"] = w
        u = rng.normal((cout,))
        v = rng.normal((cin * k * k,))
        st[p + "weight_u"] = u / u.norm()
        st[p + "weight_v"] = v / v.norm()
    else:
        st[p + "weight"] = w


def _bn(st: State, p: str, rng: _Rng, c: int):
    # non-trivial affine / running stats so that parity checks exercise them
    st[p + "weight"] = 1.0 + rng.uniform((c,), 0.2)
    st[p + "bias"] = rng.uniform((c,), 0.2)
    st[p + "running_mean"] = rng.uniform((c,), 0.1)
    st[p + "running_var"] = 1.0 + rng.uniform((c,), 0.1)
    st[p + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _prelu(st: State, p: str):
    st[p + "weight"] = torch.full((1,), 0.25)


def generator_state(seed: int, n_blocks: int = 16, nf: int = 64, nf_last: int = 256,
                    scales: Sequence[int] = (2,), use_sn: bool = True, n_suffix: int = 0,
                    in_ch: int = 3) -> State:
    """Keys of Generator (model_generator.py:23-63) wrapped in ``n_suffix`` GeneratorSuffix."""
    rng = _Rng(seed)
    base = "base." * n_suffix
    st: State = {}
    # outer-first registration order: base.* , upscale.* for each suffix level
    p = base
    _conv(st, p + "first_layers.0.", rng, in_ch, nf, 9, True)
    _prelu(st, p + "first_layers.1.")
    for i in range(n_blocks):
        q = f"{p}block_list.{i}.layers."
        _conv(st, q + "0.", rng, nf, nf, 3, True)
        _bn(st, q + "1.", rng, nf)
        _prelu(st, q + "2.")
        _conv(st, q + "3.", rng, nf, nf, 3, True)
        _bn(st, q + "4.", rng, nf)
    _conv(st, p + "block_list_end.0.", rng, nf, nf, 3, True)
    _bn(st, p + "block_list_end.1.", rng, nf)
    cin = nf
    for s, r in enumerate(scales):
        _conv(st, f"{p}upscale.{s}.0.", rng, cin, nf_last, 3, use_sn)
        _prelu(st, f"{p}upscale.{s}.2.")
        cin = nf_last // (r * r)
    _conv(st, p + "end.0.", rng, cin, in_ch, 3, use_sn)
    for lvl in range(n_suffix, 0, -1):
        q = "base." * (lvl - 1)
        _conv(st, q + "upscale.0.", rng, nf_last // 4, nf_last, 3, True)
        _prelu(st, q + "upscale.2.")
    return st


def progressive_state(seed: int, n_blocks: int = 2, nf: int = 64, n_suffix: int = 1, in_ch: int = 3) -> State:
    """Keys of model_generator_progressive.GeneratorSuffix chained ``n_suffix`` times over a
    GeneratorProgresiveBase (each stage wraps the previous stage's ``beginning``; channels shrink x1/4)."""
    rng = _Rng(seed)
    st: State = {}
    p = "beginning." + "0." * n_suffix               # the base network sits n_suffix Sequentials deep
    _conv(st, p + "first_layers.0.", rng, in_ch, nf, 9, False)
    _prelu(st, p + "first_layers.1.")
    for i in range(n_blocks):
        q = f"{p}block_list.{i}.layers."
        _conv(st, q + "0.", rng, nf, nf, 3, False)
        _bn(st, q + "1.", rng, nf)
        _prelu(st, q + "2.")
        _conv(st, q + "3.", rng, nf, nf, 3, False)
        _bn(st, q + "4.", rng, nf)
    _conv(st, p + "block_list_end.0.", rng, nf, nf, 3, False)
    _bn(st, p + "block_list_end.1.", rng, nf)
    c = nf
    for j in range(1, n_suffix + 1):                   # stage j (1 = innermost): conv c->c, shuffle to c/4
        q = "beginning." + "0." * (n_suffix - j)
        _conv(st, q + "1.", rng, c, c, 3, False)
        _prelu(st, q + "3.")
        c //= 4
    _conv(st, "end.0.", rng, c, in_ch, 3, False)
    return st


def discriminator_state(seed: int, input_shape=(3, 96, 96),
                        features: Sequence[int] = (64, 64, 128, 128, 256, 256, 512, 512),
                        strides: Sequence[int] = (1, 2, 1, 2, 1, 2, 1, 2)) -> State:
    """Keys of Discriminator (model_discriminator.py:19-53)."""
    rng = _Rng(seed)
    st: State = {}
    c, h, w = input_shape
    _conv(st, "conv.0.", rng, c, features[0], 3, True)
    for k in range(1, len(features)):
        q = f"conv.2.{k - 1}.layers."
        _conv(st, q + "0.", rng, features[k - 1], features[k], 3, True)
        _bn(st, q + "1.", rng, features[k])
    fc_in = w * h * features[-1] // (4 ** (sum(strides) - len(strides)))
    fc_mid = features[-1] * 2
    st["fc.0.weight"] = rng.uniform((fc_mid, fc_in), 1.0 / math.sqrt(fc_in))
    st["fc.0.bias"] = rng.uniform((fc_mid,), 1.0 / math.sqrt(fc_in))
    st["fc.2.weight"] = rng.uniform((1, fc_mid), 1.0 / math.sqrt(fc_mid))
    st["fc.2.bias"] = rng.uniform((1,), 1.0 / math.sqrt(fc_mid))
    return st


def vgg_state(seed: int, mask: int) -> State:
    """Keys of MaskedVGG.layers (model_content_extractor.py:43): ``layers.{idx}.weight/bias`` for
    the convs of torchvision vgg19().features[:k].  He-normal weights (pretrained weights cannot be
    downloaded; BASELINE.json prescribes random-init weights)."""
    rng = _Rng(seed)
    st: State = {}
    k = vgg_kept_positions(mask)[-1]
    for i, (kind, cin, cout) in enumerate(vgg19_feature_layers()[:k]):
        if kind == "conv":
            st[f"layers.{i}.weight"] = rng.normal((cout, cin, 3, 3), math.sqrt(2.0 / (cin * 9)))
            st[f"layers.{i}.bias"] = rng.uniform((cout,), 0.05)
    return st


def clone_state(st: State) -> State:
    return {k: v.clone() for k, v in st.items()}


def synthetic_hr(seed: int, batch: int, size: int, ch: int = 3) -> torch.Tensor:
    """HR patches ~ U[-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((batch, ch, size, size), generator=g) * 2 - 1
