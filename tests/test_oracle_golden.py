"""CPU: the oracle (oracle/srgan_oracle.py) must reproduce the golden vectors that
oracle/validate_against_reference.py recorded from the real reference modules."""
import os

import pytest
import math

import torch

from oracle import srgan_oracle as O
from oracle import state_factory as S


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("n_suffix", [0, 1, 2])
def test_generator_matches_reference(golden_dir, n_suffix):
    g = _load(golden_dir, f"generator_suffix{n_suffix}")
    st = S.generator_state(g["seed"], n_blocks=2, n_suffix=n_suffix)
    names = O.trainable_names(st)
    leaf = O._leaf(st, names)
    y = O.generator_forward(leaf, g["x"], training=True)
    assert y.shape == g["y"].shape == (2, 3, 8 * 2 ** (1 + n_suffix), 8 * 2 ** (1 + n_suffix))
    assert O.rel_l2(y.detach(), g["y"]) < 2e-5
    grads = dict(zip(names, torch.autograd.grad((y * g["gy"]).sum(), [leaf[k] for k in names])))
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] > 1e-3 * max(g["grad_norms"].values()):
            assert O.rel_l2(grads[k], ref) < 1e-4, k
    assert O.rel_l2(O.generator_forward(st, g["x"], training=False), g["y_eval"]) < 2e-5


@pytest.mark.parametrize("n_suffix", [1, 2])
def test_progressive_generator_matches_reference(golden_dir, n_suffix):
    """model_generator_progressive.py (the reference's older chained-suffix design)."""
    g = _load(golden_dir, f"progressive_suffix{n_suffix}")
    st = S.progressive_state(g["seed"], n_blocks=2, nf=64, n_suffix=n_suffix)
    names = O.trainable_names(st)
    leaf = O._leaf(st, names)
    y = O.progressive_forward(leaf, g["x"], training=True)
    assert y.shape == g["y"].shape == (2, 3, 8 * 2 ** n_suffix, 8 * 2 ** n_suffix)
    assert O.rel_l2(y.detach(), g["y"]) < 2e-5
    grads = dict(zip(names, torch.autograd.grad((y * g["gy"]).sum(), [leaf[k] for k in names])))
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] > 1e-3 * max(g["grad_norms"].values()):
            assert O.rel_l2(grads[k], ref) < 1e-4, k
    assert O.rel_l2(O.progressive_forward(st, g["x"], training=False), g["y_eval"]) < 2e-5


def test_discriminator_matches_reference(golden_dir):
    g = _load(golden_dir, "discriminator")
    st = S.discriminator_state(g["seed"], g["shape"], g["features"], g["strides"])
    names = O.trainable_names(st)
    leaf = O._leaf(st, names)
    x = g["x"].clone().requires_grad_(True)
    out = O.discriminator_forward(leaf, x, g["strides"], True)
    assert out.shape == (4, 1)
    assert O.rel_l2(out.detach(), g["out"]) < 2e-5
    grads = torch.autograd.grad(O.bce(out.view(-1), 0.9), [x] + [leaf[k] for k in names])
    assert O.rel_l2(grads[0], g["dx"]) < 1e-4
    by_name = dict(zip(names, grads[1:]))
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] > 1e-3 * max(g["grad_norms"].values()):
            assert O.rel_l2(by_name[k], ref) < 1e-4, k


@pytest.mark.parametrize("mask", [0b00010, 0b10000, 0b01111])
def test_masked_vgg_matches_reference(golden_dir, mask):
    g = _load(golden_dir, f"vgg_mask{mask:05b}")
    st = S.vgg_state(g["seed"], mask)
    x = g["x"].clone().requires_grad_(True)
    feat = O.masked_vgg_forward(st, x, mask)
    assert feat.shape == g["features"].shape
    assert O.rel_l2(feat.detach(), g["features"]) < 2e-5
    loss = torch.mean((g["target"] - feat) ** 2)
    assert abs(float(loss) - g["loss"]) < 1e-5 * abs(g["loss"])
    (dx,) = torch.autograd.grad(loss, [x])
    assert O.rel_l2(dx, g["dx"]) < 1e-4


def test_lr_from_hr_matches_reference(golden_dir):
    g = _load(golden_dir, "lr_from_hr")
    assert torch.equal(O.lr_from_hr(g["hr"], (4, 4)), g["lr"])
    assert float(g["lr"].abs().max()) <= 1.0


@pytest.mark.parametrize("n_steps", [1, 2])
def test_train_step_matches_reference_train_loop(golden_dir, n_steps):
    g = _load(golden_dir, f"train_step{n_steps}")
    seed = g["seed"]
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
    d_st = S.discriminator_state(seed + 1, g["shape"], g["features"], g["strides"])
    v_st = S.vgg_state(seed + 2, g["mask"])
    og = O.AdamState(O.trainable_names(g_st), g["lr"])
    od = O.AdamState(O.trainable_names(d_st), g["lr"])
    for i in range(n_steps):
        hr = S.synthetic_hr(seed + 10 + i, g["B"], g["HR"])
        out = O.train_step(g_st, d_st, v_st, hr, O.lr_from_hr(hr, (g["LR"], g["LR"])),
                           d_strides=g["strides"], vgg_mask=g["mask"], opt_g=og, opt_d=od)
        tol = 2e-5 if i == 0 else 5e-4
        assert abs(out["err_d"] - g["err_d"][i]) <= tol * abs(g["err_d"][i])
        assert abs(out["err_g_adv"] - g["err_g_adv"][i]) <= tol * abs(g["err_g_adv"][i])
        assert abs(out["err_g_cont"] - g["err_g_cont"][i]) <= 5e-4 * abs(g["err_g_cont"][i])
    if n_steps == 1:
        for k, ref in g["post_g_sample"].items():
            assert float(((g_st[k] - ref).abs() > 0.1 * g["lr"]).float().mean()) <= 5e-3, k
        for k, ref in g["post_d_sample"].items():
            assert float(((d_st[k] - ref).abs() > 0.1 * g["lr"]).float().mean()) <= 5e-3, k


@pytest.mark.parametrize("case", ["on_lr", "no_adv", "adv_only", "identity_x10"])
def test_train_step_branches_match_reference_train_loop(golden_dir, case):
    """content_loss_on_lr (train.py:41-50, 95-97), zero-weight skips (train.py:56,86,94,106) and identity()
    x10 (config.py:156-163) against the losses of the unmodified reference train_loop."""
    g = _load(golden_dir, "train_step_branches")
    c = g["cases"][case]
    wg, wd, wc, kind = c["weights"]
    seed = g["seed"]
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
    d_st = S.discriminator_state(seed + 1, g["shape"], g["features"], g["strides"])
    d_init = S.clone_state(d_st)
    v_st = S.vgg_state(seed + 2, g["mask"])
    og = O.AdamState(O.trainable_names(g_st), g["lr"])
    od = O.AdamState(O.trainable_names(d_st), g["lr"])
    for i in range(2):
        hr = S.synthetic_hr(seed + 10 + i, g["B"], g["HR"])
        hr2 = S.synthetic_hr(seed + 30 + i, g["B"], g["HR"]) if c["content_loss_on_lr"] else None
        out = O.train_step(g_st, d_st, v_st, hr, O.lr_from_hr(hr, (g["LR"], g["LR"])), d_strides=g["strides"],
                           vgg_mask=g["mask"], opt_g=og, opt_d=od, w_adv_g=wg, w_adv_d=wd, w_cont=wc,
                           content_loss_on_lr=c["content_loss_on_lr"], hr2=hr2, cont_kind=kind or "features")
        for key in ("err_d", "err_g_adv", "err_g_cont"):
            want = c[key][i]
            assert abs(out[key] - want) <= 5e-4 * abs(want), (case, i, key, out[key], want)
    if not wd and not wg:
        assert all(torch.equal(d_st[k], d_init[k]) for k in d_init)      # D untouched when its branches are off


def test_bce_restatement_matches_torch():
    p = torch.tensor([1e-9, 0.3, 0.999999, 1.0, 0.0])
    for t in (0.0, 0.9, 1.0):
        want = torch.nn.BCELoss()(p, torch.full_like(p, t))
        assert abs(float(O.bce(p, t)) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))


def test_adam_restatement_matches_torch():
    torch.manual_seed(0)
    w = torch.randn(50)
    p = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.9, 0.999))
    st = {"w": w.clone()}
    mine = O.AdamState(["w"], 1e-3)
    for i in range(3):
        g = torch.randn(50)
        p.grad = g.clone()
        opt.step()
        mine.step(st, {"w": g})
    assert O.rel_l2(st["w"], p.detach()) < 1e-6


def test_ssim_restatement_properties():
    """The oracle's SSIM (published definition; the reference has none): 1 for identical images, symmetric,
    lower for a noisier copy, and equal to the closed form for constant images."""
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, 3, 24, 24, generator=g) * 2 - 1
    assert torch.allclose(O.ssim(a, a), torch.ones(2, dtype=torch.float64))
    b1 = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(-1, 1)
    b2 = (a + 0.3 * torch.randn(a.shape, generator=g)).clamp(-1, 1)
    assert torch.allclose(O.ssim(a, b1), O.ssim(b1, a))
    assert (O.ssim(a, b1) > O.ssim(a, b2)).all() and (O.ssim(a, b2) < 0.9).all()
    x, y = torch.full((1, 1, 16, 16), 0.5), torch.full((1, 1, 16, 16), -0.25)
    c1, c2 = (0.01 * 2) ** 2, (0.03 * 2) ** 2
    want = (2 * 0.5 * -0.25 + c1) * c2 / ((0.25 + 0.0625 + c1) * c2)
    assert abs(float(O.ssim(x, y)) - want) < 1e-9
    assert abs(float(O.psnr_per_image(x, y)) - 10 * math.log10(4 / 0.75 ** 2)) < 1e-9
