"""Pin the oracle to the real reference and (re)generate tests/golden.  TEST INFRASTRUCTURE ONLY.

Runs ONLY in the dev container (needs /root/reference, which does not travel to the GPU box):

    python -m oracle.validate_against_reference            # check + write tests/golden/*.pt

For every case the unmodified reference modules (and, for the step case, the unmodified
``train.train_loop``) are executed on seeded weights/inputs; the oracle must reproduce outputs,
parameter gradients, buffer updates and post-step weights to <= 2e-5 relative L2 (fp32 round-off
from a different but equivalent op order).  The reference's outputs are then stored as the golden
vectors; the weights are not stored, only their seed (oracle/state_factory.py).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import srgan_oracle as O
from . import state_factory as S

REF = os.environ.get("SISR_REFERENCE", "/root/reference")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TOL = 2e-5


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torchvision.models as tvm
    if not getattr(tvm, "_sisr_patched", False):
        orig = tvm.vgg19
        tvm.vgg19 = lambda pretrained=True: orig(weights=None)   # no network: random init
        tvm._sisr_patched = True
    import model_generator, model_discriminator, model_content_extractor  # noqa
    return model_generator, model_discriminator, model_content_extractor


def _check(name, got, want, tol=TOL, floor=0.0):
    """relative L2; ``floor`` bounds the denominator for quantities that are analytically zero
    (e.g. the bias gradient of a conv that feeds a train-mode BatchNorm)."""
    got, want = got.detach(), want.detach()
    err = float((got.double() - want.double()).norm() / max(float(want.double().norm()), floor, 1e-30))
    status = "ok" if err <= tol else "FAIL"
    print(f"  {name:48s} rel_l2={err:.2e} {status}")
    if err > tol:
        raise SystemExit(f"oracle deviates from the reference at {name}")


def _grads(module):
    return {k: (p.grad.clone() if p.grad is not None else None) for k, p in module.named_parameters()}


def case_generator(mg, n_suffix, seed, gold):
    name = f"generator_suffix{n_suffix}"
    print(name)
    st = S.generator_state(seed, n_blocks=2, n_suffix=n_suffix)
    net = mg.Generator(2, 64, 256, [2], use_sn=True)
    for _ in range(n_suffix):
        net = mg.GeneratorSuffix(net)
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    x = S.synthetic_hr(seed + 1, 2, 8)
    gy = torch.randn(2, 3, 8 * 2 ** (1 + n_suffix), 8 * 2 ** (1 + n_suffix),
                     generator=torch.Generator().manual_seed(seed + 2))
    net.train()
    y = net(x)
    (y * gy).sum().backward()
    ref_grads = _grads(net)
    ref_sd = {k: v.clone() for k, v in net.state_dict().items()}

    mine = S.clone_state(st)
    names = O.trainable_names(mine)
    leaf = O._leaf(mine, names)
    y2 = O.generator_forward(leaf, x, training=True)
    g2 = torch.autograd.grad((y2 * gy).sum(), [leaf[k] for k in names])
    _check("forward", y2, y)
    floor = 1e-3 * max(float(v.norm()) for v in ref_grads.values())
    for k, g in zip(names, g2):
        _check("grad " + k, g, ref_grads[k], tol=1e-4, floor=floor)
    for k in mine:
        if k.endswith(("weight_u", "weight_v", "running_mean", "running_var")):
            _check("buffer " + k, mine[k], ref_sd[k])
    # eval-mode forward (visualisation.py:17-26)
    net.eval()
    with torch.no_grad():
        ye = net(x)
    _check("eval forward", O.generator_forward(mine, x, training=False), ye)
    gold[name] = {"seed": seed, "n_suffix": n_suffix, "x": x, "gy": gy, "y": y.detach(),
                  "y_eval": ye,
                  "grad_norms": {k: float(v.norm()) for k, v in ref_grads.items()},
                  "grads": {k: ref_grads[k] for k in names if ref_grads[k].numel() <= 40000 and
                            ("block_list.1." in k or "first_layers" in k or "upscale" in k
                             or "end" in k)}}


def case_discriminator(md, seed, gold):
    print("discriminator")
    shape, feats, strides = (3, 16, 16), [64, 64, 128, 128], [1, 2, 1, 2]
    st = S.discriminator_state(seed, shape, feats, strides)
    net = md.Discriminator(shape, feats, strides)
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    x = S.synthetic_hr(seed + 1, 4, 16).requires_grad_(True)
    net.train()
    out = net(x)
    loss = O.bce(out.view(-1), 0.9)
    loss.backward()
    ref_grads = _grads(net)
    mine = S.clone_state(st)
    names = O.trainable_names(mine)
    leaf = O._leaf(mine, names)
    x2 = x.detach().clone().requires_grad_(True)
    out2 = O.discriminator_forward(leaf, x2, strides, True)
    g2 = torch.autograd.grad(O.bce(out2.view(-1), 0.9), [leaf[k] for k in names] + [x2])
    _check("forward", out2, out)
    _check("bce vs nn.BCELoss", O.bce(out.view(-1), 0.9),
           torch.nn.BCELoss()(out.view(-1), torch.full((4,), 0.9)))
    floor = 1e-3 * max(float(v.norm()) for v in ref_grads.values())
    for k, g in zip(names, g2[:-1]):
        _check("grad " + k, g, ref_grads[k], tol=1e-4, floor=floor)
    _check("grad input", g2[-1], x.grad, tol=1e-4)
    gold["discriminator"] = {"seed": seed, "shape": shape, "features": feats, "strides": strides,
                             "x": x.detach(), "out": out.detach(), "dx": x.grad.clone(),
                             "grad_norms": {k: float(v.norm()) for k, v in ref_grads.items()},
                             "grads": {k: v for k, v in ref_grads.items() if "fc.2" in k or
                                       "conv.0" in k or "conv.2.2.layers.1" in k}}


def case_vgg(mc, mask, seed, size, gold):
    name = f"vgg_mask{mask:05b}"
    print(name)
    st = S.vgg_state(seed, mask)
    net = mc.MaskedVGG(mask)
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    x = S.synthetic_hr(seed + 1, 2, size).requires_grad_(True)
    feat = net(x)
    assert feat.shape[1] == mc.get_size(x, mask)
    target = torch.randn(feat.shape, generator=torch.Generator().manual_seed(seed + 2)) * 0.1
    loss = torch.mean((target - feat) ** 2)
    loss.backward()
    x2 = x.detach().clone().requires_grad_(True)
    feat2 = O.masked_vgg_forward(st, x2, mask)
    (dx2,) = torch.autograd.grad(torch.mean((target - feat2) ** 2), [x2])
    _check("features", feat2, feat)
    _check("grad input", dx2, x.grad, tol=1e-4)
    gold[name] = {"seed": seed, "mask": mask, "x": x.detach(), "target": target,
                  "features": feat.detach(), "dx": x.grad.clone(), "loss": float(loss)}


def case_train_step(mg, md, mc, seed, gold, n_steps=2):
    """Unmodified train.train_loop with an injected synthetic ``config``."""
    print("train_step (reference train.train_loop, unmodified)")
    B, HR, LRs = 4, 16, 4
    shape, feats, strides = (3, HR, HR), [64, 64, 128, 128], [1, 2, 1, 2]
    mask = 0b00010
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
    d_st = S.discriminator_state(seed + 1, shape, feats, strides)
    v_st = S.vgg_state(seed + 2, mask)
    batches = [S.synthetic_hr(seed + 10 + i, B, HR) for i in range(n_steps + 1)]

    net_g = mg.GeneratorSuffix(mg.Generator(2, 64, 256, [2], use_sn=True))
    net_d = md.Discriminator(shape, feats, strides)
    ext = mc.MaskedVGG(mask)
    torch.nn.Module.load_state_dict(net_g, S.clone_state(g_st), strict=True)
    torch.nn.Module.load_state_dict(net_d, S.clone_state(d_st), strict=True)
    torch.nn.Module.load_state_dict(ext, S.clone_state(v_st), strict=True)

    import utils as ref_utils
    ref_utils.save_curr_vis = lambda *a, **k: None
    lr = 1e-3  # larger than config.py:38 so that two steps move the weights measurably
    cfg = types.ModuleType("config")
    opt_g = torch.optim.Adam(net_g.parameters(), lr=lr, betas=(.9, .999))
    opt_d = torch.optim.Adam(net_d.parameters(), lr=lr, betas=(.9, .999))
    cfg.__dict__.update(dict(
        torch=torch, np=np, random=__import__("random"), utils=ref_utils,
        device=torch.device("cpu"), net_g=net_g, net_d=net_d, net_content_extractor=ext,
        criterion=torch.nn.BCELoss(), optimizerG=opt_g, optimizerD=opt_d,
        schedulerG=torch.optim.lr_scheduler.LambdaLR(opt_g, lambda it: 1),
        schedulerD=torch.optim.lr_scheduler.LambdaLR(opt_d, lambda it: 1),
        real_label=torch.full((B,), 1.0), real_label_reduced=torch.full((B,), .9),
        fake_label=torch.full((B,), .0),
        loss_weight_adv_g=lambda e: 5e-2, loss_weight_adv_d=lambda e: 1.0,
        loss_weight_cont=lambda e: (1.0, ext),
        dataloader_hr=[(b, None) for b in batches], image_size_lr=(3, LRs, LRs),
        n_batch=n_steps + 1, num_epochs=1, starting_epoch=0,
        dis_list_old=[], dis_list_old_len=1000, dis_list_old_freq=1, dis_list_old_ratio=.01,
        dis_list_old_cpu=True, dis_list_old_save=False, content_loss_on_lr=False,
        plot_first=False, plot_training=False, plot_usr=False, test_hr=None, test_lr=None,
        write_root="/tmp/", identity=None))
    sys.modules["config"] = cfg
    sys.modules.pop("train", None)
    import train as ref_train
    d_losses, g_losses, c_losses, _ = ref_train.train_loop()

    og = O.AdamState(O.trainable_names(g_st), lr)
    od = O.AdamState(O.trainable_names(d_st), lr)
    mine = []
    for i in range(n_steps):
        lr_img = O.lr_from_hr(batches[i], (LRs, LRs))
        mine.append(O.train_step(g_st, d_st, v_st, batches[i], lr_img, d_strides=strides,
                                 vgg_mask=mask, opt_g=og, opt_d=od))
    # Biases of convs that feed a train-mode BatchNorm have an analytically zero gradient; Adam
    # turns the fp32 round-off noise in them into +-lr updates, so they (and, through round-off,
    # later-step losses at the 1e-5 level) are not comparable between two correct implementations.
    def noise_driven(k):
        return k.endswith(("layers.0.bias", "layers.3.bias", "block_list_end.0.bias")) and "conv.0." not in k
    for i in range(n_steps):
        tol = TOL if i == 0 else 5e-4
        _check(f"step{i} err_d", torch.tensor(mine[i]["err_d"]), torch.tensor(d_losses[i]), tol)
        _check(f"step{i} err_g_adv", torch.tensor(mine[i]["err_g_adv"]), torch.tensor(g_losses[i]), tol)
        _check(f"step{i} err_g_cont", torch.tensor(mine[i]["err_g_cont"]), torch.tensor(c_losses[i]),
               max(tol, 1e-4))
    # Early Adam steps are sign-like (|update| ~ lr whatever |g|), so an element whose gradient is
    # at round-off level may move the other way; parity of the updated weights is therefore
    # "all but a vanishing fraction of elements agree to 10 % of lr".
    ref_g, ref_d = net_g.state_dict(), net_d.state_dict()
    # (after the first step only: from the second step on the two runs are legitimately on
    # slightly different trajectories and only the losses are compared)
    for tag, mine_st, ref_st in (("G", g_st, ref_g), ("D", d_st, ref_d)):
        if n_steps > 1:
            break
        for k in mine_st:
            if not mine_st[k].dtype.is_floating_point or noise_driven(k):
                continue
            if k.endswith(("weight_u", "weight_v", "running_mean", "running_var")):
                _check(f"post-step {tag} {k}", mine_st[k], ref_st[k], tol=2e-3)
                continue
            frac = float(((mine_st[k] - ref_st[k]).abs() > 0.1 * lr).float().mean())
            status = "ok" if frac <= 5e-3 else "FAIL"
            print(f"  post-step {tag} {k:40s} moved-differently fraction={frac:.1e} {status}")
            if frac > 5e-3:
                raise SystemExit("oracle deviates from the reference at post-step " + k)
    gold[f"train_step{n_steps}"] = {
        "seed": seed, "B": B, "HR": HR, "LR": LRs, "shape": shape, "features": feats,
        "strides": strides, "mask": mask, "lr": lr, "n_steps": n_steps,
        "err_d": d_losses, "err_g_adv": g_losses, "err_g_cont": c_losses,
        "post_g_norms": {k: float(v.double().norm()) for k, v in ref_g.items()
                         if v.dtype.is_floating_point},
        "post_d_norms": {k: float(v.double().norm()) for k, v in ref_d.items()
                         if v.dtype.is_floating_point},
        "post_g_sample": {k: ref_g[k].clone() for k in ("upscale.0.bias", "base.end.0.bias",
                                                        "base.first_layers.1.weight",
                                                        "base.block_list.0.layers.1.weight")},
        "post_d_sample": {k: ref_d[k].clone() for k in ("fc.2.weight", "fc.2.bias", "conv.0.bias")},
    }


def case_train_step_branches(mg, md, mc, seed, gold):
    """Unmodified train.train_loop in the "unsupervised" mode (content_loss_on_lr=True, config.py:24:
    HR swap train.py:41-50, identity content loss x100 on the re-downsampled fake train.py:95-97, adversarial
    weight 5e-3) and with epoch-scheduled weights that switch branches off (train.py:56,86,94,106)."""
    from . import ref_harness as R
    import model_content_extractor as mce
    B, HR, LRs = 4, 16, 4
    shape, feats, strides, mask, lr = (3, HR, HR), [64, 64, 128, 128], [1, 2, 1, 2], 0b00010, 1e-3
    cases = {
        # name: (content_loss_on_lr, weights (adv_g, adv_d, cont, kind))
        "on_lr": (True, (5e-3, 1.0, 100.0, "identity")),
        "no_adv": (False, (0, 0, 1.0, "features")),          # pre-training on the content loss only
        "adv_only": (False, (5e-2, 1.0, 0, None)),
        "identity_x10": (False, (5e-2, 1.0, 10.0, "identity")),
    }
    out = {}
    for name, (on_lr, (wg, wd, wc, kind)) in cases.items():
        print("train_step branch", name)
        g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
        d_st = S.discriminator_state(seed + 1, shape, feats, strides)
        v_st = S.vgg_state(seed + 2, mask)
        hrs = [S.synthetic_hr(seed + 10 + i, B, HR) for i in range(3)]
        hr2s = [S.synthetic_hr(seed + 30 + i, B, HR) for i in range(3)]
        net_g = mg.GeneratorSuffix(mg.Generator(2, 64, 256, [2], use_sn=True))
        net_d = md.Discriminator(shape, feats, strides)
        ext = mc.MaskedVGG(mask)
        for net, st in ((net_g, g_st), (net_d, d_st), (ext, v_st)):
            torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
        ident = mce.identity()
        lw = (lambda e, w=wg: w, lambda e, w=wd: w,
              lambda e, w=wc, k=kind: (w, {"features": ext, "identity": ident, None: None}[k]))
        batches = [((h, None), (h2, None)) for h, h2 in zip(hrs, hr2s)] if on_lr else hrs
        d_l, g_l, c_l, _, _ = R.run_train_loop(net_g, net_d, ext, batches, lr=lr, lr_size=LRs,
                                               content_loss_on_lr=on_lr, loss_weights=lw, identity=ident)
        og = O.AdamState(O.trainable_names(g_st), lr)
        od = O.AdamState(O.trainable_names(d_st), lr)
        for i in range(2):
            lr_img = O.lr_from_hr(hrs[i], (LRs, LRs))
            r = O.train_step(g_st, d_st, v_st, hrs[i], lr_img, d_strides=strides, vgg_mask=mask, opt_g=og,
                             opt_d=od, w_adv_g=wg, w_adv_d=wd, w_cont=wc, content_loss_on_lr=on_lr,
                             hr2=hr2s[i] if on_lr else None, cont_kind=kind or "features")
            tol = TOL if i == 0 else 5e-4
            for key, want in (("err_d", d_l[i]), ("err_g_adv", g_l[i]), ("err_g_cont", c_l[i])):
                if want == 0.0:
                    assert r[key] == 0.0, (name, key)
                else:
                    _check(f"{name} step{i} {key}", torch.tensor(r[key]), torch.tensor(want), max(tol, 1e-4))
        # a skipped D update leaves the discriminator untouched (weights AND spectral-norm / BN buffers)
        if not wd and not wg:
            ref_d = net_d.state_dict()
            init = S.discriminator_state(seed + 1, shape, feats, strides)
            assert all(torch.equal(ref_d[k], init[k]) for k in init), "reference moved D with zero weights"
            assert all(torch.equal(d_st[k], init[k]) for k in init), "oracle moved D with zero weights"
        out[name] = {"content_loss_on_lr": on_lr, "weights": (wg, wd, wc, kind),
                     "err_d": d_l, "err_g_adv": g_l, "err_g_cont": c_l}
    gold["train_step_branches"] = {"seed": seed, "B": B, "HR": HR, "LR": LRs, "shape": shape, "features": feats,
                                   "strides": strides, "mask": mask, "lr": lr, "cases": out}


def case_lr_from_hr(seed, gold):
    print("lr_from_hr")
    import utils as ref_utils
    hr = S.synthetic_hr(seed, 2, 16)
    want = ref_utils.lr_from_hr(hr, (4, 4))
    _check("bicubic+clamp", O.lr_from_hr(hr, (4, 4)), want, tol=0.0)
    gold["lr_from_hr"] = {"seed": seed, "hr": hr, "lr": want}


def case_progressive(n_suffix, seed, gold):
    """model_generator_progressive.py (older design, not imported by config/train): forward, all
    parameter gradients and the eval-mode forward of the chained GeneratorSuffix stages."""
    import model_generator_progressive as mp
    name = f"progressive_suffix{n_suffix}"
    print(name)
    st = S.progressive_state(seed, n_blocks=2, nf=64, n_suffix=n_suffix)
    net = mp.GeneratorSuffix(mp.GeneratorProgresiveBase(2, 64), 64)
    nf = 16
    for _ in range(n_suffix - 1):
        net = mp.GeneratorSuffix(net.beginning, nf)
        nf //= 4
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    x = S.synthetic_hr(seed + 1, 2, 8)
    side = 8 * 2 ** n_suffix
    gy = torch.randn(2, 3, side, side, generator=torch.Generator().manual_seed(seed + 2))
    net.train()
    y = net(x)
    (y * gy).sum().backward()
    ref_grads = _grads(net)
    ref_sd = {k: v.clone() for k, v in net.state_dict().items()}
    mine = S.clone_state(st)
    names = O.trainable_names(mine)
    leaf = O._leaf(mine, names)
    y2 = O.progressive_forward(leaf, x, training=True)
    g2 = torch.autograd.grad((y2 * gy).sum(), [leaf[k] for k in names])
    _check("forward", y2, y)
    floor = 1e-3 * max(float(v.norm()) for v in ref_grads.values())
    for k, g in zip(names, g2):
        _check("grad " + k, g, ref_grads[k], tol=1e-4, floor=floor)
    for k in mine:
        if k.endswith(("running_mean", "running_var")):
            _check("buffer " + k, mine[k], ref_sd[k])
    net.eval()
    with torch.no_grad():
        ye = net(x)
    _check("eval forward", O.progressive_forward(mine, x, training=False), ye)
    gold[name] = {"seed": seed, "n_suffix": n_suffix, "x": x, "gy": gy, "y": y.detach(), "y_eval": ye,
                  "grad_norms": {k: float(v.norm()) for k, v in ref_grads.items()},
                  "grads": {k: ref_grads[k] for k in names if ref_grads[k].numel() <= 40000 and
                            ("block_list.1." in k or "first_layers" in k or "beginning.1" in k
                             or "beginning.3" in k or "end" in k)}}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    mg, md, mc = _import_reference()
    gold = {}
    case_generator(mg, 0, 100, gold)
    case_generator(mg, 1, 110, gold)
    case_generator(mg, 2, 120, gold)
    case_discriminator(md, 200, gold)
    case_vgg(mc, 0b00010, 300, 16, gold)
    case_vgg(mc, 0b10000, 310, 32, gold)
    case_vgg(mc, 0b01111, 320, 16, gold)
    case_lr_from_hr(400, gold)
    case_progressive(1, 600, gold)
    case_progressive(2, 610, gold)
    case_train_step(mg, md, mc, 500, gold, n_steps=1)
    case_train_step(mg, md, mc, 500, gold, n_steps=2)
    case_train_step_branches(mg, md, mc, 520, gold)
    os.makedirs(GOLD, exist_ok=True)
    for name, blob in gold.items():
        torch.save(blob, os.path.join(GOLD, name + ".pt"))
    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"wrote {len(gold)} golden files, {total / 1e6:.2f} MB, torch {torch.__version__}")


if __name__ == "__main__":
    main()
