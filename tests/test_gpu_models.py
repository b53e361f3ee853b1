"""GPU parity of the drop-in modules and of the whole training step against the CPU oracle on the
seeds of the golden fixtures (which pin the oracle to the real reference).

Tolerances
  * forward activations / outputs vs the fp32 reference (golden): relative L2 <= 1e-2 and
    super-resolved images >= 50 dB PSNR (north_star);
  * per-layer gradients vs fp32: tests/test_gpu_ops.py (relative L2 <= 1e-2 per operator);
  * END-TO-END gradients are judged against the oracle run with bf16 *storage emulation* (fp32
    arithmetic, tensors rounded where the CUDA path stores bf16): relative L2 <= 8e-2 for the
    single networks (outputs / losses <= 1e-2).  What is left is the fp32 summation order: a
    value that lands within fp32 round-off of a bf16 rounding boundary is stored one bf16 ulp
    apart, and downstream LeakyReLU(0.01)/PReLU/max-pool decisions amplify that.  For the deep
    chains (VGG54, the whole G-through-D-and-VGG step) the bound is therefore self-calibrated:
    the oracle is run twice more with a 1e-6 relative jitter before each bf16 rounding
    (``O.jitter_before_rounding``); the distance between those two CPU runs is the noise floor
    of ANY correct bf16-storage implementation, and the CUDA path must be within 1.5x of it.  Against the pure fp32
    reference an end-to-end gradient of a ReLU-family network with bf16 activations cannot agree
    to 1e-2: ~1 % of the units sit within bf16 round-off of zero and take the other branch
    (LeakyReLU slope 0.01, PReLU 0.25, max-pool routing), which moves the relative L2 by 10-50 %
    while the cosine stays 0.85-0.99.  That bound is asserted too, on the reference's own values.
"""
import os

import pytest
import torch

from oracle import srgan_oracle as O
from oracle import state_factory as S

pytestmark = pytest.mark.gpu


def rel(a, b):
    return O.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())


def cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _generator(m, n_suffix, st):
    net = m.Generator(2, 64, 256, [2], use_sn=True)
    for _ in range(n_suffix):
        net = m.GeneratorSuffix(net)
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    return net.cuda()


@pytest.mark.parametrize("n_suffix", [0, 1, 2])
def test_generator_vs_reference_golden(cuda, golden_dir, n_suffix):
    import sisr_b200 as m
    g = _load(golden_dir, f"generator_suffix{n_suffix}")
    st = S.generator_state(g["seed"], n_blocks=2, n_suffix=n_suffix)
    net = _generator(m, n_suffix, st)
    net.train()
    y = net(g["x"].cuda())
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    assert O.psnr(y.detach().cpu(), g["y"]) >= 50.0
    assert rel(y, g["y"]) < 1e-2
    (y * g["gy"].cuda()).sum().backward()
    grads = {k: p.grad for k, p in net.named_parameters()}
    top = max(g["grad_norms"].values())
    # (a) vs the real reference's gradients: direction and magnitude
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] > 1e-2 * top:       # skip analytically-zero gradients (bias before BN)
            assert cos(grads[k], ref) > 0.85, k
    for k, n_ref in g["grad_norms"].items():
        if n_ref > 1e-2 * top:
            # a PReLU slope gradient is ONE number summed over every unit of its layer: with bf16 storage the
            # oracle itself lands 3-23 % away from the fp32 value (7 jittered CPU runs, round 2), tensors do not
            assert abs(float(grads[k].norm()) - n_ref) < (0.30 if grads[k].numel() == 1 else 0.15) * n_ref, k
    # (b) vs the oracle with bf16 storage emulation: tight
    emu = S.generator_state(g["seed"], n_blocks=2, n_suffix=n_suffix)
    names = O.trainable_names(emu)
    leaf = O._leaf(emu, names)
    with O.emulate_bf16_storage():
        y_emu = O.generator_forward(leaf, g["x"], training=True)
        g_emu = dict(zip(names, torch.autograd.grad((y_emu * g["gy"]).sum(), [leaf[k] for k in names])))
    assert rel(y, y_emu) < 1e-2
    top_e = max(float(v.norm()) for v in g_emu.values())
    for k, r in g_emu.items():
        if float(r.norm()) > 1e-2 * top_e:
            # scalars (PReLU slopes): one number summed over a whole layer, see (a); which side of 8 % it lands on
            # changes with the accumulation order inside the conv kernel (igemm_th 0.078, igemm_pm 0.082)
            assert rel(grads[k], r) < (0.25 if r.numel() == 1 else 8e-2), k
    # buffers advanced exactly once (spectral-norm u/v, BN running stats)
    ref_state = S.generator_state(g["seed"], n_blocks=2, n_suffix=n_suffix)
    O.generator_forward(ref_state, g["x"], training=True)
    sd = net.state_dict()
    for k in ref_state:
        if k.endswith(("weight_u", "weight_v")):
            assert rel(sd[k], ref_state[k]) < 1e-4, k
        if k.endswith(("running_mean", "running_var")):
            assert rel(sd[k], ref_state[k]) < 1e-2, k
    net.eval()
    with torch.no_grad():
        ye = net(g["x"].cuda())
    assert O.psnr(ye.cpu(), g["y_eval"]) >= 50.0


@pytest.mark.parametrize("n_suffix", [1, 2])
def test_progressive_generator_vs_reference_golden(cuda, golden_dir, n_suffix):
    """model_generator_progressive.py drop-in (GeneratorProgresiveBase + chained GeneratorSuffix stages,
    64 -> 16 channels) against the real reference's outputs and gradients."""
    from sisr_b200 import model_generator_progressive as mp
    g = _load(golden_dir, f"progressive_suffix{n_suffix}")
    st = S.progressive_state(g["seed"], n_blocks=2, nf=64, n_suffix=n_suffix)
    net = mp.GeneratorSuffix(mp.GeneratorProgresiveBase(2, 64), 64)
    nf = 16
    for _ in range(n_suffix - 1):
        net = mp.GeneratorSuffix(net.beginning, nf)
        nf //= 4
    assert set(net.state_dict()) == set(st)
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    net = net.cuda().train()
    y = net(g["x"].cuda())
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    assert rel(y, g["y"]) < 1e-2
    (y * g["gy"].cuda()).sum().backward()
    grads = {k: p.grad for k, p in net.named_parameters()}
    top = max(g["grad_norms"].values())
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] > 1e-2 * top:
            assert cos(grads[k], ref) > 0.85, k
    for k, n_ref in g["grad_norms"].items():
        if n_ref > 1e-2 * top:
            assert abs(float(grads[k].norm()) - n_ref) < (0.30 if grads[k].numel() == 1 else 0.15) * n_ref, k
    emu = S.progressive_state(g["seed"], n_blocks=2, nf=64, n_suffix=n_suffix)
    names = O.trainable_names(emu)
    leaf = O._leaf(emu, names)
    with O.emulate_bf16_storage():
        y_emu = O.progressive_forward(leaf, g["x"], training=True)
        g_emu = dict(zip(names, torch.autograd.grad((y_emu * g["gy"]).sum(), [leaf[k] for k in names])))
    assert rel(y, y_emu) < 1e-2
    top_e = max(float(v.norm()) for v in g_emu.values())
    for k, r in g_emu.items():
        if float(r.norm()) > 1e-2 * top_e:
            assert rel(grads[k], r) < 8e-2, k
    net.eval()
    with torch.no_grad():
        ye = net(g["x"].cuda())
    assert rel(ye, g["y_eval"]) < 1e-2
    if n_suffix == 2:
        # the reference's own smoke test chains one more stage down to 4 channels
        # (model_generator_progressive.py:67-86: g3 = GeneratorSuffix(g2.beginning, n_features=4))
        st3 = S.progressive_state(g["seed"], n_blocks=2, nf=64, n_suffix=3)
        net.train()
        g3 = mp.GeneratorSuffix(net.beginning, 4).cuda().train()
        assert set(g3.state_dict()) == set(st3)
        torch.nn.Module.load_state_dict(g3, S.clone_state(st3), strict=True)
        y3 = g3(g["x"].cuda())
        assert y3.shape == (g["x"].shape[0], 3, 8 * 8, 8 * 8)
        names = O.trainable_names(st3)
        leaf = O._leaf(S.clone_state(st3), names)
        y3_ref = O.progressive_forward(leaf, g["x"], training=True)
        assert rel(y3, y3_ref) < 1e-2, rel(y3, y3_ref)
        gy3 = torch.randn(y3_ref.shape, generator=torch.Generator().manual_seed(9))
        want = dict(zip(names, torch.autograd.grad((y3_ref * gy3).sum(), [leaf[k] for k in names])))
        (y3 * gy3.cuda()).sum().backward()
        got = {k: p.grad for k, p in g3.named_parameters()}
        for k in ("end.0.weight", "end.0.bias", "beginning.1.weight", "beginning.3.weight"):
            assert got[k] is not None and got[k].shape == want[k].shape, k
            assert cos(got[k], want[k]) > 0.9, (k, cos(got[k], want[k]))
        with pytest.raises(NotImplementedError):
            mp.GeneratorSuffix(net.beginning, 12)


def test_discriminator_vs_reference_golden(cuda, golden_dir):
    import sisr_b200 as m
    g = _load(golden_dir, "discriminator")
    st = S.discriminator_state(g["seed"], g["shape"], g["features"], g["strides"])
    net = m.Discriminator(g["shape"], g["features"], g["strides"])
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    net = net.cuda().train()
    x = g["x"].cuda().requires_grad_(True)
    out = net(x)
    assert out.shape == (4, 1)
    assert rel(out, g["out"]) < 1e-2
    from sisr_b200 import ops
    loss, _ = ops.bce_loss(out.view(-1), 0.9)
    loss.backward()
    assert cos(x.grad, g["dx"]) > 0.9
    grads = {k: p.grad for k, p in net.named_parameters()}
    top = max(g["grad_norms"].values())
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] > 1e-2 * top:
            assert cos(grads[k], ref) > 0.85, k
    def emulated(jitter_seed=None):
        emu = S.discriminator_state(g["seed"], g["shape"], g["features"], g["strides"])
        names = O.trainable_names(emu)
        leaf = O._leaf(emu, names)
        xe = g["x"].clone().requires_grad_(True)
        with O.emulate_bf16_storage():
            jit = O.jitter_before_rounding(1e-6, jitter_seed) if jitter_seed is not None else None
            if jit:
                jit.__enter__()
            out_e = O.discriminator_forward(leaf, xe, g["strides"], True)
            ge = torch.autograd.grad(O.bce(out_e.view(-1), 0.9), [xe] + [leaf[k] for k in names])
            if jit:
                jit.__exit__()
        return names, out_e, ge
    names, out_e, ge = emulated()
    _, _, j1 = emulated(1)
    _, _, j2 = emulated(2)
    assert rel(out, out_e) < 1e-2
    # LeakyReLU(0.01) makes every unit that flips branch under bf16 round-off a 100x change, so the
    # end-to-end gradient is compared against the distance between two jittered oracle runs.
    # Measured over 6 seeds (tools/gpu_diag_dnoise.py, gpurun_out/t14): CUDA-vs-oracle 0.064-0.101,
    # oracle-vs-jittered-oracle 0.086-0.116; a single pair of jittered runs scatters by +-30 %, hence
    # the floor of 0.15 under the 1.5x rule.
    assert rel(x.grad, ge[0]) < max(0.15, 1.5 * rel(j1[0], j2[0])), rel(j1[0], j2[0])
    top_e = max(float(v.norm()) for v in ge[1:])
    for i, (k, r) in enumerate(zip(names, ge[1:]), 1):
        if float(r.norm()) > 1e-2 * top_e:
            assert rel(grads[k], r) < max(0.15, 1.5 * rel(j1[i], j2[i])), (k, rel(j1[i], j2[i]))


@pytest.mark.parametrize("mask", [0b00010, 0b10000, 0b01111])
def test_masked_vgg_vs_reference_golden(cuda, golden_dir, mask):
    import sisr_b200 as m
    from sisr_b200 import ops
    g = _load(golden_dir, f"vgg_mask{mask:05b}")
    net = m.MaskedVGG(mask)
    torch.nn.Module.load_state_dict(net, S.vgg_state(g["seed"], mask), strict=True)
    net = net.cuda()
    x = g["x"].cuda().requires_grad_(True)
    feat = net(x)
    assert feat.shape == g["features"].shape
    assert rel(feat, g["features"]) < 1e-2
    loss = ops.mse_loss(g["target"].cuda(), feat)
    assert abs(float(loss.detach()) - g["loss"]) < 2e-2 * g["loss"]
    loss.backward()
    assert cos(x.grad, g["dx"]) > 0.9
    xe = g["x"].clone().requires_grad_(True)
    with O.emulate_bf16_storage():
        fe = O.masked_vgg_forward(S.vgg_state(g["seed"], mask), xe, mask)
        (dxe,) = torch.autograd.grad(torch.mean((g["target"] - fe) ** 2), [xe])
    assert rel(feat, fe) < 1e-2

    def jittered(seed):
        xj = g["x"].clone().requires_grad_(True)
        with O.emulate_bf16_storage(), O.jitter_before_rounding(1e-6, seed):
            fj = O.masked_vgg_forward(S.vgg_state(g["seed"], mask), xj, mask)
            (dxj,) = torch.autograd.grad(torch.mean((g["target"] - fj) ** 2), [xj])
        return dxj
    floor = rel(jittered(1), jittered(2))
    assert rel(x.grad, dxe) < max(8e-2, 1.5 * floor), (rel(x.grad, dxe), floor)


def _build_step(m, seed, shape, feats, strides, mask, lr):
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
    d_st = S.discriminator_state(seed + 1, shape, feats, strides)
    v_st = S.vgg_state(seed + 2, mask)
    net_g = m.GeneratorSuffix(m.Generator(2, 64, 256, [2], use_sn=True))
    net_d = m.Discriminator(shape, feats, strides)
    ext = m.MaskedVGG(mask)
    torch.nn.Module.load_state_dict(net_g, S.clone_state(g_st), strict=True)
    torch.nn.Module.load_state_dict(net_d, S.clone_state(d_st), strict=True)
    torch.nn.Module.load_state_dict(ext, S.clone_state(v_st), strict=True)
    tr = m.SRGANTrainer(net_g.cuda(), net_d.cuda(), ext.cuda(), m.StepConfig(lr=lr, use_replay=False))
    return tr, (g_st, d_st, v_st)


@pytest.mark.parametrize("graph", [False, True])
def test_train_step_vs_reference_train_loop(cuda, golden_dir, graph):
    """Losses of the unmodified reference train.train_loop (golden) vs the CUDA step."""
    import sisr_b200 as m
    g = _load(golden_dir, "train_step2")
    tr, _ = _build_step(m, g["seed"], g["shape"], g["features"], g["strides"], g["mask"], g["lr"])
    hrs = [S.synthetic_hr(g["seed"] + 10 + i, g["B"], g["HR"]) for i in range(2)]
    lrs = [O.lr_from_hr(h, (g["LR"], g["LR"])) for h in hrs]
    outs = []
    if graph:
        # capture with throw-away weights state: rebuild afterwards so step 0 starts from the seed
        tr.capture(hrs[0].cuda(), lrs[0].cuda(), warmup=1)
        tr2, _ = _build_step(m, g["seed"], g["shape"], g["features"], g["strides"], g["mask"], g["lr"])
        for dst, src in ((tr.net_g, tr2.net_g), (tr.net_d, tr2.net_d)):
            with torch.no_grad():
                for (k, a), (_, b) in zip(dst.state_dict().items(), src.state_dict().items()):
                    a.copy_(b)
        for opt in (tr.opt_g, tr.opt_d):
            for stt in opt.state.values():
                stt["exp_avg"].zero_(); stt["exp_avg_sq"].zero_()
            for step_t, _ in opt._dev_state.values():
                step_t.zero_()
        for i in range(2):
            o = tr.replay(hrs[i].cuda(), lrs[i].cuda())
            outs.append({k: float(o[k]) for k in ("err_d", "err_g_adv", "err_g_cont")})
    else:
        for i in range(2):
            o = tr.step(hrs[i].cuda(), lrs[i].cuda())
            outs.append({k: float(o[k]) for k in ("err_d", "err_g_adv", "err_g_cont")})
    for i in range(2):
        # step 1 starts from weights that an lr = 1e-3 sign-like Adam update has moved: six jittered CPU oracle
        # runs of this very case scatter by +-3 % on err_g_adv there, and the fp32 L2 reductions of the small
        # weight gradients add a run-to-run component on the GPU (measured up to 6 %)
        tol = 2e-2 if i == 0 else 8e-2
        for k in ("err_d", "err_g_adv", "err_g_cont"):
            assert abs(outs[i][k] - g[k][i]) < tol * abs(g[k][i]), (i, k, outs[i][k], g[k][i])


def test_train_step_gradients_and_update_vs_oracle(cuda):
    """One step at a slightly larger size: every parameter gradient the optimisers see, the fake
    batch (PSNR) and the direction of the Adam update against the oracle."""
    import sisr_b200 as m
    seed, shape, feats, strides, mask, lr = 700, (3, 32, 32), [64, 64, 128, 128, 256, 256], \
        [1, 2, 1, 2, 1, 2], 0b00110, 1e-3
    tr, (g_st, d_st, v_st) = _build_step(m, seed, shape, feats, strides, mask, lr)
    hr = S.synthetic_hr(seed + 5, 4, 32)
    lr_img = O.lr_from_hr(hr, (8, 8))
    g_before = {k: v.detach().clone() for k, v in tr.net_g.state_dict().items()}
    out = tr.step(hr.cuda(), lr_img.cuda())
    fp32_states = (S.clone_state(g_st), S.clone_state(d_st), S.clone_state(v_st))
    ref32 = O.train_step(*fp32_states, hr, lr_img, d_strides=strides, vgg_mask=mask,
                         opt_g=O.AdamState(O.trainable_names(g_st), lr),
                         opt_d=O.AdamState(O.trainable_names(d_st), lr))
    assert O.psnr(out["fake"].float().cpu(), ref32["fake"]) >= 50.0
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(out[k]) - ref32[k]) < 2e-2 * abs(ref32[k]), k
    with O.emulate_bf16_storage():
        ref = O.train_step(g_st, d_st, v_st, hr, lr_img, d_strides=strides, vgg_mask=mask,
                           opt_g=O.AdamState(O.trainable_names(g_st), lr),
                           opt_d=O.AdamState(O.trainable_names(d_st), lr))
    # noise floor: two jittered CPU runs of the same step
    def jittered(sd):
        gj = S.generator_state(seed, n_blocks=2, n_suffix=1)
        dj = S.discriminator_state(seed + 1, shape, feats, strides)
        with O.emulate_bf16_storage(), O.jitter_before_rounding(1e-6, sd):
            res = O.train_step(gj, dj, S.vgg_state(seed + 2, mask), hr, lr_img, d_strides=strides,
                               vgg_mask=mask, opt_g=O.AdamState(O.trainable_names(gj), lr),
                               opt_d=O.AdamState(O.trainable_names(dj), lr))
        res["g_after"] = gj
        return res
    j1, j2 = jittered(1), jittered(2)
    # err_g_adv is evaluated AFTER the sign-like first Adam step of D (lr 1e-3), so it inherits the
    # round-off sensitivity of that update: 1 % or twice the worst deviation of a jittered run
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        tol = max(1e-2 * abs(ref[k]), 2.0 * max(abs(j1[k] - ref[k]), abs(j2[k] - ref[k])))
        assert abs(float(out[k]) - ref[k]) < tol, (k, float(out[k]), ref[k], j1[k], j2[k])
    g_grads = {k: p.grad for k, p in tr.net_g.named_parameters()}
    d_grads = {k: p.grad for k, p in tr.net_d.named_parameters()}
    # D .grad holds nothing from the G step (its weight gradient is skipped there), i.e. exactly
    # the D-update gradient that the reference would have used
    for tag, mine, key in (("G", g_grads, "g_grads"), ("D", d_grads, "d_grads")):
        top = max(float(v.norm()) for v in ref[key].values())
        for k, r in ref[key].items():
            if float(r.norm()) > 1e-2 * top:
                floor = rel(j1[key][k], j2[key][k])
                err = rel(mine[k], r)
                # scalars (PReLU slopes): one pair of jittered runs is a one-sample estimate of the noise
                bound = max(1e-1, 2.5 * floor) if r.numel() == 1 else max(8e-2, 1.5 * floor)
                assert err < bound, (tag, k, err, floor)
                assert cos(mine[k], r) > 0.8, (tag, k)
    # Adam moved the trainable weights the same way (the first step is sign-like, so the measure is
    # the share of elements that moved in the same direction, against the share on which two
    # jittered CPU runs of the same step agree with each other)
    g_init = S.generator_state(seed, n_blocks=2, n_suffix=1)
    for k in ("base.end.0.weight_orig", "upscale.0.weight_orig", "base.block_list.1.layers.3.weight_orig"):
        mine = tr.net_g.state_dict()[k].cpu() - g_before[k].cpu()
        want = g_st[k] - g_init[k]
        same = float((torch.sign(mine) == torch.sign(want)).float().mean())
        floor = float((torch.sign(j1["g_after"][k] - g_init[k]) ==
                       torch.sign(j2["g_after"][k] - g_init[k])).float().mean())
        assert same > min(0.9, floor - 0.05), (k, same, floor)


def test_frozen_prefix_trains_only_the_suffix(cuda):
    """config 4: x2 weights wrapped by GeneratorSuffix(freeze_prefix=True, ...) (model_generator.py:161-184)."""
    import sisr_b200 as m
    g1 = m.Generator(2, 64, 256, [2], use_sn=True)
    g2 = m.GeneratorSuffix(g1, freeze_prefix=True, freeze_upscale=True, freeze_end=True).cuda()
    before = {k: p.detach().clone() for k, p in g2.named_parameters()}
    opt = m.Adam(g2.parameters(), lr=0.1, betas=(.9, .999))
    res = g2(torch.rand(4, 3, 8, 8, device="cuda") * 2 - 1)
    assert res.shape == (4, 3, 32, 32)
    (res - torch.zeros_like(res)).pow(2).sum().backward()
    opt.step()
    for k, p in g2.named_parameters():
        changed = bool((p.detach() != before[k]).any())
        assert changed == (not k.startswith("base.")), k


def test_checkpoint_resume_and_torch_adam_interchange(cuda):
    """utils._save dictionary (utils.py:107-114): resume continues exactly where the run stopped, and
    the optimizer state loads into torch.optim.Adam (and back)."""
    import sisr_b200 as m
    seed, shape, feats, strides, mask, lr = 710, (3, 16, 16), [64, 64, 128, 128], [1, 2, 1, 2], 0b00010, 1e-3
    hrs = [S.synthetic_hr(seed + 10 + i, 4, 16) for i in range(3)]
    lrs = [O.lr_from_hr(h, (4, 4)) for h in hrs]
    tr, _ = _build_step(m, seed, shape, feats, strides, mask, lr)
    for i in range(2):
        tr.step(hrs[i].cuda(), lrs[i].cuda())
    ckpt = tr.checkpoint(epoch=7)
    assert set(ckpt) == {"epoch", "net_g", "net_d", "opti_g", "opti_d", "dis_list"}
    ckpt = {k: (S.clone_state(v) if k.startswith("net_") else v) for k, v in ckpt.items()}
    import copy
    ckpt["opti_g"], ckpt["opti_d"] = copy.deepcopy(ckpt["opti_g"]), copy.deepcopy(ckpt["opti_d"])
    want = tr.step(hrs[2].cuda(), lrs[2].cuda())
    tr2, _ = _build_step(m, seed, shape, feats, strides, mask, lr)               # untrained weights
    ckpt["net_g"] = {"module." + k: v for k, v in ckpt["net_g"].items()}         # saved under DataParallel
    assert tr2.restore(ckpt) == 7
    got = tr2.step(hrs[2].cuda(), lrs[2].cuda())
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(got[k]) - float(want[k])) < 5e-3 * abs(float(want[k])), k
    # optimizer interchange: same keys / shapes as torch.optim.Adam, step counter included
    params = [torch.nn.Parameter(p.detach().clone()) for p in tr.net_d.parameters()]
    ref = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
    ref.load_state_dict(ckpt["opti_d"])
    st = ref.state[params[0]]
    assert int(st["step"]) == 2 and st["exp_avg"].shape == params[0].shape
    sched = torch.optim.lr_scheduler.LambdaLR(ref, lr_lambda=lambda it: 0.9 ** it)   # config.py:170-180
    for p in params:
        p.grad = torch.ones_like(p)
    ref.step()                                     # the reference's optimizer must STEP on our state
    sched.step()
    assert int(ref.state[params[0]]["step"]) == 3
    tr2.opt_d.load_state_dict(ref.state_dict())
    assert int(tr2.opt_d._dev_state[0][0].item()) == 3
    assert abs(tr2.opt_d.param_groups[0]["lr"] - lr) < 1e-12      # schedule restarts from initial_lr
    out = tr2.step(hrs[2].cuda(), lrs[2].cuda())                     # ... and ours on the reference's state
    assert all(torch.isfinite(out[k]).all() for k in ("err_d", "err_g_adv", "err_g_cont"))
    assert int(tr2.opt_d._dev_state[0][0].item()) == 4
    # the replay list is exported as the reference keeps it (fp32, CPU) and comes back as bf16 GPU tensors
    tr2.cfg.use_replay = True
    tr2.train_iteration(hrs[0].cuda(), lrs[0].cuda())
    ck2 = tr2.checkpoint()
    assert ck2["dis_list"][0].dtype == torch.float32 and not ck2["dis_list"][0].is_cuda
    st_d = S.discriminator_state(seed + 1, shape, feats, strides)
    assert O.discriminator_forward(st_d, ck2["dis_list"][0], strides, True).shape == (4, 1)   # fp32 net accepts it
    tr2.restore(ck2)
    assert tr2.dis_list_old[0].dtype == torch.bfloat16 and tr2.dis_list_old[0].is_cuda


def test_experience_replay_step_vs_oracle(cuda):
    """train.py:59-71,144-146: replayed old fakes add one discriminator pass each, their BCE terms are
    SUMMED into the D loss; the replay list lives on the GPU in bf16."""
    import sisr_b200 as m
    seed, shape, feats, strides, mask, lr = 720, (3, 16, 16), [64, 64, 128, 128], [1, 2, 1, 2], 0b00010, 1e-3
    tr, (g_st, d_st, v_st) = _build_step(m, seed, shape, feats, strides, mask, lr)
    hr = S.synthetic_hr(seed + 10, 4, 16)
    lr_img = O.lr_from_hr(hr, (4, 4))
    old = [(S.synthetic_hr(seed + 20 + i, 4, 16) * 0.5).to(torch.bfloat16).float() for i in range(2)]
    out = tr.step(hr.cuda(), lr_img.cuda(), [o.cuda() for o in old])
    ref = O.train_step(g_st, d_st, v_st, hr, lr_img, d_strides=strides, vgg_mask=mask,
                       opt_g=O.AdamState(O.trainable_names(g_st), lr),
                       opt_d=O.AdamState(O.trainable_names(d_st), lr), old_fakes=old)
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(out[k]) - ref[k]) < 3e-2 * abs(ref[k]), (k, float(out[k]), ref[k])
    assert abs(float(out["d_g_z1"]) - ref["d_g_z1"]) < 3e-2 * abs(ref["d_g_z1"])      # sum over 3 fake batches
    # bookkeeping of train_iteration: the list grows by one bf16 GPU tensor per step
    tr.cfg.use_replay = True
    tr.train_iteration(hr.cuda(), lr_img.cuda())
    assert len(tr.dis_list_old) == 1 and tr.dis_list_old[0].dtype == torch.bfloat16 and tr.dis_list_old[0].is_cuda


def test_progressive_x8_step_at_256_vs_oracle(cuda):
    """BASELINE.json configs[4]: x8 = GeneratorSuffix(GeneratorSuffix(Generator)) on 256x256 HR patches
    (LR 32x32), D at (3,256,256) (fc_in = 131 072), MaskedVGG54 - one full step at batch 2 against the
    CPU oracle; also configs[3]: the trunk frozen (freeze_prefix/upscale/end) in a second trainer."""
    import sisr_b200 as m
    seed, shape, feats, strides, mask, lr = 730, (3, 256, 256), [64, 64, 128, 128, 256, 256, 512, 512], \
        [1, 2, 1, 2, 1, 2, 1, 2], 0b10000, 1e-4
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=2)
    d_st = S.discriminator_state(seed + 1, shape, feats, strides)
    v_st = S.vgg_state(seed + 2, mask)
    net_g = m.GeneratorSuffix(m.GeneratorSuffix(m.Generator(2, 64, 256, [2], use_sn=True)))
    net_d = m.Discriminator(shape, feats, strides)
    assert net_d.fc_in == 131072
    ext = m.MaskedVGG(mask)
    torch.nn.Module.load_state_dict(net_g, S.clone_state(g_st), strict=True)
    torch.nn.Module.load_state_dict(net_d, S.clone_state(d_st), strict=True)
    torch.nn.Module.load_state_dict(ext, S.clone_state(v_st), strict=True)
    tr = m.SRGANTrainer(net_g.cuda(), net_d.cuda(), ext.cuda(), m.StepConfig(lr=lr, use_replay=False))
    hr = S.synthetic_hr(seed + 5, 2, 256)
    lr_img = O.lr_from_hr(hr, (32, 32))
    out = tr.step(hr.cuda(), lr_img.cuda())
    assert out["fake"].shape == (2, 3, 256, 256)
    ref = O.train_step(g_st, d_st, v_st, hr, lr_img, d_strides=strides, vgg_mask=mask,
                       opt_g=O.AdamState(O.trainable_names(g_st), lr),
                       opt_d=O.AdamState(O.trainable_names(d_st), lr))
    assert O.psnr(out["fake"].float().cpu(), ref["fake"]) >= 50.0
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(out[k]) - ref[k]) < 2e-2 * abs(ref[k]), (k, float(out[k]), ref[k])
    # configs[3]: x2 weights wrapped and frozen - only the new suffix stage trains
    g1 = m.Generator(2, 64, 256, [2], use_sn=True)
    g2 = m.GeneratorSuffix(g1, freeze_prefix=True, freeze_upscale=True, freeze_end=True).cuda()
    trainable = [k for k, p in g2.named_parameters() if p.requires_grad]
    assert sorted(trainable) == ["upscale.0.bias", "upscale.0.weight_orig", "upscale.2.weight"]
    d2 = m.Discriminator((3, 32, 32), [64, 64], [1, 2]).cuda()
    tr2 = m.SRGANTrainer(g2, d2, m.MaskedVGG(0b00010).cuda(), m.StepConfig(lr=1e-3, use_replay=False))
    before = {k: p.detach().clone() for k, p in g2.named_parameters()}
    hr2 = S.synthetic_hr(seed + 6, 4, 32)
    tr2.step(hr2.cuda(), O.lr_from_hr(hr2, (8, 8)).cuda())
    for k, p in g2.named_parameters():
        assert bool((p.detach() != before[k]).any()) == (not k.startswith("base.")), k


def test_host_feed_pipeline_matches_direct_replay(cuda, golden_dir):
    """HostFeed + replay_from_feed (H2D of batch i+1 overlapping step i, LR made on the device) gives the
    same three losses, step by step, as replay() on device-resident tensors."""
    import sisr_b200 as m
    from sisr_b200.train import HostFeed
    g = _load(golden_dir, "train_step2")
    hrs = [S.synthetic_hr(g["seed"] + 40 + i, g["B"], g["HR"]) for i in range(4)]
    keys = ("err_d", "err_g_adv", "err_g_cont")
    res = []
    for mode in ("direct", "feed"):
        # lr 1e-5 (config.py:38), not the golden's 1e-3: the test is about the inputs the pipeline delivers.  Two
        # runs of the same step differ in the last bits of a few fp32 reductions (atomics), Adam's first updates
        # are sign-like, and at lr = 1e-3 that alone moved err_g_adv of the SECOND step by 0.1 ... 17 % between
        # two identical runs (tools/determinism_check.py, profiles/r2_determinism.txt)
        tr, _ = _build_step(m, g["seed"], g["shape"], g["features"], g["strides"], g["mask"], 1e-5)
        lr0 = m.lr_from_hr(hrs[0].cuda(), (g["LR"], g["LR"]))
        tr.capture(hrs[0].cuda(), lr0, warmup=1)
        outs = []
        if mode == "direct":
            for h in hrs:
                hd = h.cuda()
                o = tr.replay(hd, m.lr_from_hr(hd, (g["LR"], g["LR"])))
                outs.append([float(o[k]) for k in keys])
        else:
            pinned = [h.pin_memory() for h in hrs]
            feed = HostFeed(tuple(hrs[0].shape), torch.device("cuda"))
            feed.submit(pinned[0])
            for i in range(len(hrs)):
                if i + 1 < len(hrs):
                    feed.submit(pinned[i + 1])
                o = tr.replay_from_feed(feed)
                outs.append([float(o[k]) for k in keys])
            with pytest.raises(RuntimeError):
                feed.take()
        res.append(outs)
    # same kernels and inputs; fp32 atomics reorder sums and the sign-like Adam updates amplify that from step
    # to step
    for i, (a, b) in enumerate(zip(res[0], res[1])):
        for x, y in zip(a, b):
            assert abs(x - y) <= (5e-3 if i == 0 else 2e-2 if i == 1 else 6e-2) * abs(x) + 1e-6, (i, a, b)


@pytest.mark.parametrize("case", ["on_lr", "no_adv", "adv_only", "identity_x10"])
def test_step_branches_vs_reference_train_loop(cuda, golden_dir, case):
    """The step's other branches - content_loss_on_lr (HR swap for D, identity content loss x100 on the
    re-downsampled fake: the bicubic BACKWARD kernel is on the path), zero-weight skips, identity() x10 -
    against the losses of the UNMODIFIED reference train_loop (golden, oracle/validate_against_reference.py)."""
    import sisr_b200 as m
    g = _load(golden_dir, "train_step_branches")
    c = g["cases"][case]
    wg, wd, wc, kind = c["weights"]
    tr, _ = _build_step(m, g["seed"], g["shape"], g["features"], g["strides"], g["mask"], g["lr"])
    tr.cfg.content_loss_on_lr = c["content_loss_on_lr"]
    tr.cfg.loss_weight_adv_g = lambda e: wg
    tr.cfg.loss_weight_adv_d = lambda e: wd
    tr.cfg.loss_weight_cont = lambda e: (wc, kind)
    d_before = {k: v.detach().clone() for k, v in tr.net_d.state_dict().items()}
    for i in range(2):
        hr = S.synthetic_hr(g["seed"] + 10 + i, g["B"], g["HR"])
        hr2 = S.synthetic_hr(g["seed"] + 30 + i, g["B"], g["HR"]).cuda() if c["content_loss_on_lr"] else None
        out = tr.step(hr.cuda(), O.lr_from_hr(hr, (g["LR"], g["LR"])).cuda(), img_hr2=hr2)
        # second step: taken after an lr = 1e-3 sign-like Adam update of both networks (measured up to 5.2 %)
        tol = 2e-2 if i == 0 else 8e-2
        for key in ("err_d", "err_g_adv", "err_g_cont"):
            want = c[key][i]
            got = float(out[key])
            if want == 0.0:
                assert got == 0.0, (case, key, got)
            else:
                assert abs(got - want) < tol * abs(want), (case, i, key, got, want)
    if not wd and not wg:
        for k, v in tr.net_d.state_dict().items():
            assert torch.equal(v, d_before[k]), k       # no D forward at all: buffers untouched too


def test_gen_losses_schedule_matches_config():
    """config.gen_losses (config.py:124-166) restated: defaults of both modes and windowed schedules."""
    from sisr_b200.train import gen_losses
    a, b, c = gen_losses(False)
    assert (a(0), b(0), c(0)) == (5e-2, 1.0, (1.0, "features")) and a(10 ** 6) == 5e-2
    a, b, c = gen_losses(True)
    assert (a(0), b(0), c(0)) == (5e-3, 1.0, (100.0, "identity"))
    a, b, c = gen_losses(False, n_g=(2, 5), n_d=(2, 5), n_content=(1, 9), n_identity=(0, 1))
    assert [a(e) for e in range(6)] == [0, 0, 5e-2, 5e-2, 5e-2, 0]
    assert [b(e) for e in (1, 2, 5)] == [0, 1.0, 0]
    assert c(0) == (10.0, "identity") and c(1) == (1.0, "features") and c(9) == (0, None)


def test_graph_replay_with_replayed_fakes(cuda, golden_dir):
    """Training past iteration 100 stays on the CUDA-graph path: one graph per number k of replayed fakes
    (train.py:144-146), captured on first use; the graph step equals the eager step on the same inputs and
    train_iteration(graph=True) keeps the replay list and the iteration counter."""
    import sisr_b200 as m
    g = _load(golden_dir, "train_step2")
    hr = S.synthetic_hr(g["seed"] + 50, g["B"], g["HR"])
    lr_img = O.lr_from_hr(hr, (g["LR"], g["LR"]))
    olds = [(S.synthetic_hr(g["seed"] + 60 + i, g["B"], g["HR"]) * 0.5).to(torch.bfloat16) for i in range(2)]
    keys = ("err_d", "err_g_adv", "err_g_cont", "d_g_z1")
    res = {}
    for mode in ("eager", "graph"):
        tr, _ = _build_step(m, g["seed"], g["shape"], g["features"], g["strides"], g["mask"], 1e-5)
        if mode == "graph":
            tr.capture(hr.cuda(), lr_img.cuda(), warmup=0)
        outs = []
        for k in (0, 2, 1, 2):
            old = [o.cuda() for o in olds[:k]]
            if mode == "graph":
                o = tr.replay(hr.cuda(), lr_img.cuda(), old_fakes=old)
            else:
                o = tr.step(hr.cuda(), lr_img.cuda(), [t.float() for t in old])
            outs.append([float(o[x]) for x in keys])
        res[mode] = outs
        if mode == "graph":
            assert sorted(k[0] for k in tr._graphs) == [0, 1, 2]
            tr.cfg.use_replay = True
            tr.cfg.dis_list_old_ratio = 0.5
            for i in range(4):
                tr.train_iteration(hr.cuda(), lr_img.cuda(), graph=True)
            assert tr.iteration == 4 and len(tr.dis_list_old) == 4
            assert tr.dis_list_old[0].dtype == torch.bfloat16
            assert tr.dis_list_old[0].data_ptr() != tr.dis_list_old[1].data_ptr()    # snapshots, not the static output
    # same kernels, same inputs: fp32 atomics reorder sums, and every later step starts from weights that
    # an lr-sized sign-like Adam update has moved (measured 5.4e-3 on err_g_adv at the 4th step)
    for i, (a, b) in enumerate(zip(res["eager"], res["graph"])):
        for x, y in zip(a, b):
            assert abs(x - y) <= (5e-3 if i == 0 else 2e-2) * abs(x) + 1e-6, (res["eager"], res["graph"])
    # more replayed fakes -> larger summed D loss (their BCE terms are added, not averaged)
    assert res["eager"][1][0] > res["eager"][0][0]
