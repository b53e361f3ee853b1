"""CPU fp32 oracle for the SRGAN training-step hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
(``single-image-super-resolution_b200``) must never route through it.

It is an independent *functional* restatement of the reference's algorithm: every network is a
pure function of a ``state`` dict that uses the reference's ``state_dict`` key names, so the same
tensors can be fed to the reference modules, to this oracle and to the CUDA path.  All arithmetic
is plain PyTorch fp32 on the CPU (the reference's own arithmetic type).

Parity pin: ``oracle/validate_against_reference.py`` runs the real reference modules and the
unmodified ``train.train_loop`` from ``/root/reference`` against these functions (outputs, every
parameter gradient, the post-step weights) and writes the golden vectors in ``tests/golden``.
The reference itself ships no golden vectors or numerical tests.

Reference citations (file:line in /root/reference):
  model_generator.py:5-19 (BasicBlock), 23-63 (Generator), 86-101 (forward), 117-141 (suffix)
  model_discriminator.py:5-15, 18-62
  model_content_extractor.py:6-7, 33-60
  train.py:33-122 (step), 128-168 (D loss), 171-181 (G adversarial loss), 183-186 (content loss)
  config.py:141,156-159 (loss weights), 186-188 (labels), 293-294 (Adam)
Third-party algorithms restated (PyTorch is an un-pinned dependency of the reference; the
de-facto pin is torch 2.11.0): legacy ``torch.nn.utils.spectral_norm`` (one power iteration per
training forward, sigma = u^T W v, gradient through sigma with u, v constant), ``BatchNorm2d``
train mode (biased batch variance for normalisation, unbiased for the running estimate,
momentum 0.1, eps 1e-5), ``BCELoss`` (log clamped at -100), ``Adam`` (bias-corrected, eps 1e-8).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, model_discriminator.py:12,40,50
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SN_EPS = 1e-12

# torchvision VGG19 "features" layout (configuration E): index -> out channels for convs.
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
             512, 512, 512, 512, "M"]
# 1-based positions of the conv that precedes each max-pool, model_content_extractor.py:6-7
VGG_TAPS_1BASED = (3, 8, 17, 26, 35)


def vgg19_feature_layers() -> List[Tuple[str, int, int]]:
    """[(kind, cin, cout)] for torchvision's vgg19().features, kind in conv/relu/pool."""
    layers, cin = [], 3
    for v in VGG19_CFG:
        if v == "M":
            layers.append(("pool", cin, cin))
        else:
            layers.append(("conv", cin, v))
            layers.append(("relu", v, v))
            cin = v
    return layers


# --------------------------------------------------------------------------- bf16 storage emulation
class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in the forward AND the backward direction (a tensor stored as bf16 in HBM)."""

    @staticmethod
    def forward(ctx, x):
        if _JITTER[0]:
            x = x * (1 + _JITTER[0] * torch.randn(x.shape, generator=_JITTER[1]))
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundGradBF16(torch.autograd.Function):
    """Identity forward, bf16 rounding of the gradient (a gradient tensor stored as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundFwdBF16(torch.autograd.Function):
    """bf16 rounding forward, exact gradient (prepared bf16 copy of an fp32 master weight)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


_EMULATE = [False]
_JITTER = [0.0, None]


class jitter_before_rounding:
    """Inside ``emulate_bf16_storage``: multiply every activation by (1 + eps*N(0,1)) with eps at
    fp32 round-off level *before* it is rounded to bf16.  Two such runs differ from each other by
    exactly what two correct implementations with different fp32 summation orders (CPU vs tensor
    cores) differ by; the tests use that run-to-run distance as the noise floor against which the
    CUDA path's end-to-end gradients are judged."""

    def __init__(self, eps: float = 1e-6, seed: int = 0):
        self.eps, self.seed = eps, seed

    def __enter__(self):
        self.prev = list(_JITTER)
        _JITTER[0], _JITTER[1] = self.eps, torch.Generator().manual_seed(self.seed)

    def __exit__(self, *exc):
        _JITTER[0], _JITTER[1] = self.prev


class emulate_bf16_storage:
    """Context manager: the oracle keeps fp32 arithmetic but rounds to bf16 exactly where the
    CUDA path stores bf16 (activations and activation gradients between kernels, prepared weight
    copies).  ReLU-family masks and max-pool routing then agree with the CUDA path, which makes
    END-TO-END gradients comparable at the 1e-2 level; against the pure-fp32 oracle the same
    gradients differ by 10-50 % in relative L2 (cosine 0.85-0.99) because ~1 % of the units whose
    pre-activation is within bf16 round-off of zero take the other branch - a property of bf16
    storage, not of the kernels (per-layer parity vs fp32 is tested separately at 1e-2)."""

    def __enter__(self):
        self.prev = _EMULATE[0]
        _EMULATE[0] = True

    def __exit__(self, *exc):
        _EMULATE[0] = self.prev


class record_taps:
    """Context manager: the generator functions record named intermediate activations (block
    outputs, trunk end, stage outputs) into the dict it yields, with ``retain_grad`` so that a later
    ``backward`` leaves the activation gradient on them - the oracle side of the per-layer parity
    checks (the CUDA modules expose the same names through ``ops.record_taps``)."""

    def __enter__(self):
        self.prev = _TAPS[0]
        _TAPS[0] = {}
        return _TAPS[0]

    def __exit__(self, *exc):
        _TAPS[0] = self.prev


_TAPS = [None]


def _tap(name: str, x: torch.Tensor) -> torch.Tensor:
    if _TAPS[0] is not None:
        if x.requires_grad:
            x.retain_grad()
        _TAPS[0][name] = x
    return x


def _q(x):      # activation stored as bf16
    return _RoundBF16.apply(x) if _EMULATE[0] else x


def _qg(x):     # gradient stored as bf16
    return _RoundGradBF16.apply(x) if _EMULATE[0] else x


def _qw(x):     # bf16 copy of a weight
    return _RoundFwdBF16.apply(x) if _EMULATE[0] else x


# --------------------------------------------------------------------------- primitives
def sn_weight(st: State, p: str, training: bool) -> torch.Tensor:
    """Legacy spectral norm: returns W_orig / sigma and (training) updates u, v in place."""
    w = st[p + "weight_orig"]
    u, v = st[p + "weight_u"], st[p + "weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            t = torch.mv(wm.t(), u)
            v.copy_(t / t.norm().clamp_min(SN_EPS))
            s = torch.mv(wm, v)
            u.copy_(s / s.norm().clamp_min(SN_EPS))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def conv(st: State, p: str, x: torch.Tensor, stride: int, pad: int, training: bool) -> torch.Tensor:
    w = sn_weight(st, p, training) if (p + "weight_orig") in st else st[p + "weight"]
    return F.conv2d(x, _qw(w), st[p + "bias"], stride=stride, padding=pad)


def batch_norm(st: State, p: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    g, b = st[p + "weight"], st[p + "bias"]
    rm, rv = st[p + "running_mean"], st[p + "running_var"]
    if not training:
        scale = g / torch.sqrt(rv + BN_EPS)
        return x * scale[None, :, None, None] + (b - rm * scale)[None, :, None, None]
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    with torch.no_grad():
        rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
        rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var * (n / max(n - 1, 1)))
        if (p + "num_batches_tracked") in st:
            st[p + "num_batches_tracked"] += 1
    xhat = (x - mean[None, :, None, None]) / torch.sqrt(var + BN_EPS)[None, :, None, None]
    return xhat * g[None, :, None, None] + b[None, :, None, None]


def prelu(st: State, p: str, x: torch.Tensor) -> torch.Tensor:
    a = st[p + "weight"]
    return torch.clamp_min(x, 0) + a * torch.clamp_max(x, 0)


def leaky(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp_min(x, 0) + LEAKY_SLOPE * torch.clamp_max(x, 0)


def pixel_shuffle(x: torch.Tensor, r: int) -> torch.Tensor:
    b, c, h, w = x.shape
    x = x.reshape(b, c // (r * r), r, r, h, w)
    return x.permute(0, 1, 4, 2, 5, 3).reshape(b, c // (r * r), h * r, w * r)


def bce(p: torch.Tensor, target: float) -> torch.Tensor:
    """nn.BCELoss(reduction='mean') against a constant target, log clamped at -100."""
    logp = torch.log(p).clamp_min(-100.0)
    log1mp = torch.log(1 - p).clamp_min(-100.0)
    return -(target * logp + (1 - target) * log1mp).mean()


# --------------------------------------------------------------------------- generator
def _count(st: State, fmt: str) -> int:
    n = 0
    while fmt.format(n) in st:
        n += 1
    return n


def generator_forward_no_end(st: State, x: torch.Tensor, p: str = "", training: bool = True,
                             scales: Optional[Sequence[int]] = None) -> torch.Tensor:
    """model_generator.py:86-96 (Generator) / 133-136 (GeneratorSuffix), selected by key layout."""
    if any(k.startswith(p + "base.") for k in st):
        x = generator_forward_no_end(st, x, p + "base.", training, scales)
        x = _qg(conv(st, p + "upscale.0.", x, 1, 1, training))
        x = pixel_shuffle(x, 2)
        return _tap(p + "upscale", _q(prelu(st, p + "upscale.2.", x)))
    x = _qg(conv(st, p + "first_layers.0.", _q(x), 1, 4, training))
    x = _tap(p + "first_layers", _q(prelu(st, p + "first_layers.1.", x)))
    skip = x
    n_blocks = _count(st, p + "block_list.{}.layers.0.bias")
    for i in range(n_blocks):
        q = f"{p}block_list.{i}.layers."
        y = _q(conv(st, q + "0.", x, 1, 1, training))
        y = batch_norm(st, q + "1.", y, training)
        y = _q(prelu(st, q + "2.", y))
        y = _q(conv(st, q + "3.", y, 1, 1, training))
        y = batch_norm(st, q + "4.", y, training)
        x = _tap(f"{p}block_list.{i}", _q(x + y))
    x = _q(conv(st, p + "block_list_end.0.", x, 1, 1, training))
    x = batch_norm(st, p + "block_list_end.1.", x, training)
    x = _tap(p + "block_list_end", _q(x + skip))
    n_up = _count(st, p + "upscale.{}.0.bias")
    for s in range(n_up):
        q = f"{p}upscale.{s}."
        x = _qg(conv(st, q + "0.", x, 1, 1, training))
        x = pixel_shuffle(x, scales[s] if scales else 2)
        x = _tap(f"{p}upscale.{s}", _q(prelu(st, q + "2.", x)))
    return x


def _end_prefix(st: State, p: str = "") -> str:
    while any(k.startswith(p + "base.") for k in st):
        p += "base."
    return p + "end.0."


def generator_forward(st: State, x: torch.Tensor, training: bool = True,
                      scales: Optional[Sequence[int]] = None) -> torch.Tensor:
    """model_generator.py:98-101 / 138-141: trunk + upscale stages, then the shared end conv + tanh."""
    x = generator_forward_no_end(st, x, "", training, scales)
    return torch.tanh(_qg(conv(st, _end_prefix(st), x, 1, 1, training)))


# --------------------------------------------------------------------------- older progressive design
def _progressive_base(st: State, x: torch.Tensor, p: str, training: bool) -> torch.Tensor:
    """model_generator_progressive.py:21-44 (GeneratorProgresiveBase: no spectral norm, no global skip)."""
    x = _qg(conv(st, p + "first_layers.0.", _q(x), 1, 4, training))
    x = _q(prelu(st, p + "first_layers.1.", x))
    n_blocks = _count(st, p + "block_list.{}.layers.0.bias")
    for i in range(n_blocks):
        q = f"{p}block_list.{i}.layers."
        y = _q(conv(st, q + "0.", x, 1, 1, training))
        y = batch_norm(st, q + "1.", y, training)
        y = _q(prelu(st, q + "2.", y))
        y = _q(conv(st, q + "3.", y, 1, 1, training))
        y = batch_norm(st, q + "4.", y, training)
        x = _q(x + y)
    x = _q(conv(st, p + "block_list_end.0.", x, 1, 1, training))
    return _q(batch_norm(st, p + "block_list_end.1.", x, training))


def progressive_forward_no_end(st: State, x: torch.Tensor, p: str = "beginning.",
                               training: bool = True) -> torch.Tensor:
    """model_generator_progressive.py:52-56: ``beginning`` = Sequential[prefix, conv3x3 nf->nf,
    PixelShuffle(2), PReLU]; the prefix (index 0) is the base network or the ``beginning`` of the
    previous stage (a Sequential again), which is read off the key layout."""
    if (p + "0.first_layers.0.weight") in st:
        x = _progressive_base(st, x, p + "0.", training)
    else:
        x = progressive_forward_no_end(st, x, p + "0.", training)
    x = _qg(conv(st, p + "1.", x, 1, 1, training))
    x = pixel_shuffle(x, 2)
    return _q(prelu(st, p + "3.", x))


def progressive_forward(st: State, x: torch.Tensor, training: bool = True) -> torch.Tensor:
    """model_generator_progressive.py:61-64: beginning, then end = conv3x3 (nf/4 -> 3) + Tanh."""
    x = progressive_forward_no_end(st, x, "beginning.", training)
    return torch.tanh(_qg(conv(st, "end.0.", x, 1, 1, training)))


# --------------------------------------------------------------------------- discriminator
def discriminator_forward(st: State, x: torch.Tensor, strides: Sequence[int],
                          training: bool = True) -> torch.Tensor:
    """model_discriminator.py:55-62.  Output shape (B, 1), post-sigmoid."""
    b = x.shape[0]
    x = _q(leaky(_qg(conv(st, "conv.0.", _q(x), strides[0], 1, training))))
    for k in range(len(strides) - 1):
        q = f"conv.2.{k}.layers."
        x = _q(conv(st, q + "0.", x, strides[k + 1], 1, training))
        x = _q(leaky(batch_norm(st, q + "1.", x, training)))
    x = x.reshape(b, -1)  # NCHW (c, h, w) order
    # the CUDA head feeds its three GEMMs bf16 operands (weight and dh rounded in flight)
    x = leaky(_qg(F.linear(x, _qw(st["fc.0.weight"]), st["fc.0.bias"])))
    x = F.linear(x, st["fc.2.weight"], st["fc.2.bias"])
    return torch.sigmoid(x)


# --------------------------------------------------------------------------- VGG extractor
def vgg_kept_positions(mask: int) -> List[int]:
    return [VGG_TAPS_1BASED[i] for i in range(5) if mask & (1 << i)]


def masked_vgg_forward(st: State, x: torch.Tensor, mask: int) -> torch.Tensor:
    """model_content_extractor.py:51-60 including the in-place-ReLU quirk: every tap except the
    last is observed *after* the following ReLU (torchvision uses ReLU(inplace=True) and the
    reference keeps a reference to the conv output), the last tap is the raw conv output."""
    kept = vgg_kept_positions(mask)
    layers = vgg19_feature_layers()[: kept[-1]]
    taps = []
    x = _q(x)
    for i, (kind, _, _) in enumerate(layers, 1):
        if kind == "conv":
            x = F.conv2d(x, _qw(st[f"layers.{i - 1}.weight"]), st[f"layers.{i - 1}.bias"], padding=1)
            if i in kept and i == kept[-1]:
                x = _q(x)
                taps.append(x)
            else:
                x = _qg(x)
        elif kind == "relu":
            x = _q(torch.relu(x))
            if (i - 1) in kept:
                taps.append(x)
        else:
            x = F.max_pool2d(x, 2, 2)
    return torch.cat([t.reshape(t.shape[0], -1) for t in taps], dim=1)


# --------------------------------------------------------------------------- optimiser
class AdamState:
    """torch.optim.Adam(lr, betas=(.9,.999), eps=1e-8, weight_decay=0) restated (config.py:293-294)."""

    def __init__(self, names: Sequence[str], lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.names, self.lr, self.betas, self.eps = list(names), lr, betas, eps
        self.step_count = 0
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}

    @torch.no_grad()
    def step(self, st: State, grads: Dict[str, Optional[torch.Tensor]], lr_scale: float = 1.0):
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.step_count
        bc2 = 1 - b2 ** self.step_count
        for k in self.names:
            g = grads.get(k)
            if g is None:
                continue
            if k not in self.m:
                self.m[k] = torch.zeros_like(st[k])
                self.v[k] = torch.zeros_like(st[k])
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            st[k].addcdiv_(self.m[k], denom, value=-(self.lr * lr_scale) / bc1)


# --------------------------------------------------------------------------- training step
def trainable_names(st: State) -> List[str]:
    skip = ("weight_u", "weight_v", "running_mean", "running_var", "num_batches_tracked")
    return [k for k in st if not k.endswith(skip)]


def _leaf(st: State, names: Sequence[str]) -> State:
    out = dict(st)
    for k in names:
        out[k] = st[k].detach().requires_grad_(True)
    return out


def train_step(g: State, d: State, vgg: State, hr: torch.Tensor, lr_img: torch.Tensor, *,
               d_strides: Sequence[int], vgg_mask: int, opt_g: AdamState, opt_d: AdamState,
               w_adv_g: float = 5e-2, w_adv_d: float = 1.0, w_cont: float = 1.0,
               real_label: float = 1.0, real_label_reduced: float = 0.9, fake_label: float = 0.0,
               g_trainable: Optional[Sequence[str]] = None, old_fakes: Sequence[torch.Tensor] = (),
               lr_scale: float = 1.0, content_loss_on_lr: bool = False, hr2: Optional[torch.Tensor] = None,
               cont_kind: str = "features") -> Dict[str, object]:
    """One iteration of train.py:33-122 (replay list passed explicitly).

    ``g``/``d`` are updated in place (weights, SN u/v, BN running stats) exactly as the reference
    modules would be.  Returns losses, the fake batch and the gradients that the optimisers saw.
    A zero weight skips its branch (train.py:56,86,94,106).  ``content_loss_on_lr`` (train.py:41-50,
    95-97): the discriminator's real batch is ``hr2`` and the content loss compares ``lr_img`` with the
    re-downsampled fake; ``cont_kind`` "identity" = model_content_extractor.identity (pixel MSE).
    """
    g_names = list(g_trainable) if g_trainable is not None else trainable_names(g)
    d_names = trainable_names(d)
    gl = _leaf(g, g_names)
    fake = generator_forward(gl, lr_img, training=True)              # train.py:53
    real_d = hr2 if (content_loss_on_lr and hr2 is not None) else hr

    # ---- D update, train.py:56-75 and 128-168
    err_d = torch.zeros(())
    d_fake_mean, d_x, d_grads = 0.0, 0.0, {}
    if w_adv_d:
        dl = _leaf(d, d_names)
        d_real = discriminator_forward(dl, real_d, d_strides, True).view(-1)
        d_x = float(d_real.detach().mean())
        err_d = bce(d_real, real_label_reduced)
        for fk in [fake.detach(), *old_fakes]:
            d_fake = discriminator_forward(dl, fk, d_strides, True).view(-1)
            err_d = err_d + bce(d_fake, fake_label)
            d_fake_mean += float(d_fake.detach().mean())
        err_d = err_d * w_adv_d
        d_grads = dict(zip(d_names, torch.autograd.grad(err_d, [dl[k] for k in d_names])))
        # (buffers were updated in place through the shared tensors of ``d``)
        opt_d.step(d, d_grads, lr_scale)

    # ---- G update, train.py:81-108, 171-186
    err_g_adv, err_g_cont, d_g_z2, g_grads = torch.zeros(()), torch.zeros(()), 0.0, {}
    if w_adv_g:
        dl2 = _leaf(d, d_names)
        out = discriminator_forward(dl2, fake, d_strides, True).view(-1)
        d_g_z2 = float(out.detach().mean())
        err_g_adv = bce(out, real_label) * w_adv_g
    if w_cont:
        if content_loss_on_lr:
            a, b = lr_img, lr_from_hr(fake, tuple(lr_img.shape[-2:]))
        else:
            a, b = hr, fake
        if cont_kind != "identity":
            a, b = masked_vgg_forward(vgg, a, vgg_mask), masked_vgg_forward(vgg, b, vgg_mask)
        err_g_cont = torch.mean((a - b) ** 2) * w_cont
    if w_adv_g or w_cont:
        err_g = err_g_adv + err_g_cont
        g_grads = dict(zip(g_names, torch.autograd.grad(err_g, [gl[k] for k in g_names],
                                                        allow_unused=True)))
        opt_g.step(g, g_grads, lr_scale)
    return {
        "fake": fake.detach(), "err_d": float(err_d), "err_g_adv": float(err_g_adv),
        "err_g_cont": float(err_g_cont), "d_x": d_x, "d_g_z1": d_fake_mean,
        "d_g_z2": d_g_z2, "d_grads": d_grads, "g_grads": g_grads,
    }


# --------------------------------------------------------------------------- LR synthesis
def lr_from_hr(hr: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """utils.py:16-31: bicubic (align_corners=True, A=-0.75) down-sampling, clamped to [-1, 1]."""
    return F.interpolate(hr, size, mode="bicubic", align_corners=True).clamp(-1.0, 1.0)


# --------------------------------------------------------------------------- metrics
def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)


def ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> torch.Tensor:
    """Per-image SSIM (Wang et al. 2004) of NCHW batches: 11 x 11 Gaussian window (sigma 1.5), K1 = 0.01,
    K2 = 0.03, valid window positions only, mean over channels and positions.  The reference has no
    implementation (README.md:88 lists PSNR / SSIM as a todo); this is the published definition."""
    a, b = a.double(), b.double()
    n, c, _, _ = a.shape
    x = torch.arange(11, dtype=torch.float64) - 5.0
    g = torch.exp(-x * x / (2 * 1.5 * 1.5))
    g = g / g.sum()
    win = (g[:, None] * g[None, :]).expand(c, 1, 11, 11)

    def blur(t):
        return F.conv2d(t, win, groups=c)
    ma, mb = blur(a), blur(b)
    va, vb, cov = blur(a * a) - ma * ma, blur(b * b) - mb * mb, blur(a * b) - ma * mb
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2))
    return s.reshape(n, -1).mean(dim=1)


def psnr_per_image(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> torch.Tensor:
    mse = ((a.double() - b.double()) ** 2).reshape(a.shape[0], -1).mean(dim=1)
    return 10.0 * torch.log10(data_range * data_range / mse)
