"""GPU parity AT THE BENCHMARKED DEPTH (BASELINE.json configs[1]: 16 residual blocks + one suffix stage,
discriminator at 96x96 with all eight convs, MaskedVGG54) - the round-1 golden tests stop at two blocks.

What is asserted, and why in this form (measured on the CPU with the oracle alone):
  * every residual block IN ISOLATION (teacher forcing: the oracle's fp32 input of block i is fed to the
    CUDA block i) reproduces the oracle's fp32 output within rel-L2 1e-2 - the north_star per-layer bound;
  * the CUMULATIVE drift of a 33-conv chain with bf16 operands is larger than that at depth: the oracle
    itself, run in fp32 arithmetic with tensors rounded to bf16 where the CUDA path stores them, is
    5e-3 away from pure fp32 after block 0 and 1.16e-2 after block 15 (an fp32 residual stream only
    brings that to 0.96e-2: the error comes from the bf16 GEMM operands, not from the skip adds).  The
    end-to-end check is therefore: block outputs within max(1e-2, 1.5 x jitter floor) of the
    bf16-storage oracle, within 1.5e-2 of pure fp32, and the SR image >= 50 dB PSNR against pure fp32;
  * activation gradients at every block against the bf16-storage oracle, bounded by the distance between
    two jittered oracle runs (a PReLU unit that flips branch under round-off changes its gradient 4x).
"""
import contextlib

import pytest
import torch

from oracle import srgan_oracle as O
from oracle import state_factory as S

pytestmark = pytest.mark.gpu

FEATS = [64, 64, 128, 128, 256, 256, 512, 512]
STRIDES = [1, 2, 1, 2, 1, 2, 1, 2]


def rel(a, b):
    return O.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())


def nchw(t):
    """NHWC bf16 tap of the CUDA path -> NCHW fp32 on the CPU."""
    return t.detach().float().permute(0, 3, 1, 2).contiguous().cpu()


def _oracle_g(seed, x, gy, emulate, jitter=None):
    st = S.generator_state(seed, n_blocks=16, n_suffix=1)
    names = O.trainable_names(st)
    leaf = O._leaf(st, names)
    with contextlib.ExitStack() as es:
        if emulate:
            es.enter_context(O.emulate_bf16_storage())
        if jitter is not None:
            es.enter_context(O.jitter_before_rounding(1e-6, jitter))
        taps = es.enter_context(O.record_taps())
        y = O.generator_forward(leaf, x, training=True)
        (y * gy).sum().backward()
    return y.detach(), {k: (v.detach(), v.grad) for k, v in taps.items()}, {k: leaf[k].grad for k in names}


def test_generator16_every_block_vs_oracle(cuda):
    """GeneratorSuffix(Generator(16, ...)) forward + backward at B=4, LR 24x24 (model_generator.py:86-101,
    133-141) with hooks on every block output on both sides."""
    import sisr_b200 as m
    from sisr_b200 import ops
    seed = 900
    x = S.synthetic_hr(seed + 1, 4, 24)
    gy = torch.randn(4, 3, 96, 96, generator=torch.Generator().manual_seed(5))
    y32, t32, g32 = _oracle_g(seed, x, gy, False)
    ye, te, ge = _oracle_g(seed, x, gy, True)
    _, tj1, gj1 = _oracle_g(seed, x, gy, True, 1)
    _, tj2, gj2 = _oracle_g(seed, x, gy, True, 2)

    net = m.GeneratorSuffix(m.Generator(16, 64, 256, [2], use_sn=True))
    torch.nn.Module.load_state_dict(net, S.clone_state(S.generator_state(seed, n_blocks=16, n_suffix=1)), strict=True)
    net = net.cuda().train()
    with ops.record_taps() as taps:
        y = net(x.cuda())
        (y * gy.cuda()).sum().backward()
    assert set(taps) == set(t32), (sorted(taps), sorted(t32))
    assert O.psnr(y.detach().cpu(), y32) >= 50.0, O.psnr(y.detach().cpu(), y32)
    report = []
    for name in t32:
        a = nchw(taps[name])
        floor = O.rel_l2(tj1[name][0], tj2[name][0])
        e_emu, e_32 = O.rel_l2(a, te[name][0]), O.rel_l2(a, t32[name][0])
        ga = nchw(taps[name].grad)
        gfloor = O.rel_l2(tj1[name][1], tj2[name][1])
        g_emu = O.rel_l2(ga, te[name][1])
        report.append((name, e_emu, e_32, floor, g_emu, gfloor))
    for name, e_emu, e_32, floor, g_emu, gfloor in report:
        print(f"{name:24s} act vs bf16-oracle {e_emu:.4f} vs fp32 {e_32:.4f} (floor {floor:.4f})   "
              f"grad vs bf16-oracle {g_emu:.4f} (floor {gfloor:.4f})")
    for name, e_emu, e_32, floor, g_emu, gfloor in report:
        assert e_emu < max(1e-2, 1.5 * floor), (name, e_emu, floor)
        assert e_32 < 1.5e-2, (name, e_32)
        assert g_emu < max(2e-2, 1.5 * gfloor), (name, g_emu, gfloor)
    # parameter gradients: the ones that carry signal, against the bf16-storage oracle and the floor
    grads = {k: p.grad for k, p in net.named_parameters()}
    top = max(float(v.norm()) for v in ge.values())
    for k, r in ge.items():
        if float(r.norm()) > 1e-2 * top:
            floor = O.rel_l2(gj1[k], gj2[k])
            # a PReLU slope gradient is ONE number (a sum over every unit of the layer): the distance of a
            # single pair of jittered runs is a one-sample estimate of its noise, hence the wider factor
            bound = max(1e-1, 2.5 * floor) if r.numel() == 1 else max(8e-2, 1.5 * floor)
            assert rel(grads[k], r) < bound, (k, rel(grads[k], r), floor)


@pytest.mark.parametrize("block", [0, 7, 15])
def test_block_teacher_forced_vs_fp32_oracle(cuda, block):
    """One BasicBlock (model_generator.py:5-19) of the 16-block trunk fed the ORACLE's fp32 input of that
    block: output within 1e-2 of the fp32 oracle (per-layer bound of north_star); the input gradient for the
    oracle's own upstream gradient within max(2e-2, 1.5 x the jitter floor of the same block)."""
    import sisr_b200 as m
    seed = 900
    x = S.synthetic_hr(seed + 1, 4, 24)
    gy = torch.randn(4, 3, 96, 96, generator=torch.Generator().manual_seed(5))
    _, t32, _ = _oracle_g(seed, x, gy, False)
    prev = "base.first_layers" if block == 0 else f"base.block_list.{block - 1}"
    x_in, (y_ref, gy_ref) = t32[prev][0], t32[f"base.block_list.{block}"]
    st = S.generator_state(seed, n_blocks=16, n_suffix=1)
    # same initial spectral-norm vectors on both sides: each runs exactly one power iteration
    q = f"base.block_list.{block}.layers."
    sub = {k[len(q):]: v.clone() for k, v in st.items() if k.startswith(q)}
    blk = m.model_generator.BasicBlock(64)
    torch.nn.Module.load_state_dict(blk.layers, sub, strict=True)
    blk = blk.cuda().train()
    xin = x_in.cuda().requires_grad_(True)
    y = blk(xin)
    assert rel(y, y_ref) < 1e-2, rel(y, y_ref)
    y.backward(gy_ref.cuda())

    def oracle_block(emulate, jitter=None):
        s2 = {"block_list.0.layers." + k: v.clone() for k, v in sub.items()}
        xi = x_in.clone().requires_grad_(True)
        with contextlib.ExitStack() as es:
            if emulate:
                es.enter_context(O.emulate_bf16_storage())
            if jitter is not None:
                es.enter_context(O.jitter_before_rounding(1e-6, jitter))
            p = "block_list.0.layers."
            h = O._q(O.conv(s2, p + "0.", O._q(xi), 1, 1, True))
            h = O._q(O.prelu(s2, p + "2.", O.batch_norm(s2, p + "1.", h, True)))
            h = O._q(O.conv(s2, p + "3.", h, 1, 1, True))
            out = O._q(xi + O.batch_norm(s2, p + "4.", h, True))
            (g,) = torch.autograd.grad(out, [xi], gy_ref)
        return out.detach(), g
    y32, g32 = oracle_block(False)
    assert O.rel_l2(y32, y_ref) < 1e-5            # the block restated here IS the oracle's block
    _, ge = oracle_block(True)
    _, j1 = oracle_block(True, 1)
    _, j2 = oracle_block(True, 2)
    floor = O.rel_l2(j1, j2)
    err = rel(xin.grad, ge)
    print(f"block {block}: out vs fp32 {rel(y, y_ref):.4f}; dx vs bf16-oracle {err:.4f} (floor {floor:.4f}), "
          f"vs fp32 {rel(xin.grad, g32):.4f}")
    assert err < max(2e-2, 1.5 * floor), (err, floor)


def _trainer(m, seed, n_blocks, mask, lr, frozen=False, batch_shape=(3, 96, 96)):
    if frozen:
        g_st = S.generator_state(seed, n_blocks=n_blocks, n_suffix=1)
        base = m.Generator(n_blocks, 64, 256, [2], use_sn=True)
        net_g = m.GeneratorSuffix(base, freeze_prefix=True, freeze_upscale=True, freeze_end=True)
    else:
        g_st = S.generator_state(seed, n_blocks=n_blocks, n_suffix=1)
        net_g = m.GeneratorSuffix(m.Generator(n_blocks, 64, 256, [2], use_sn=True))
    d_st = S.discriminator_state(seed + 1, batch_shape, FEATS, STRIDES)
    v_st = S.vgg_state(seed + 2, mask)
    net_d = m.Discriminator(batch_shape, FEATS, STRIDES)
    ext = m.MaskedVGG(mask)
    for net, st in ((net_g, g_st), (net_d, d_st), (ext, v_st)):
        torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    tr = m.SRGANTrainer(net_g.cuda(), net_d.cuda(), ext.cuda(), m.StepConfig(lr=lr, use_replay=False))
    return tr, (g_st, d_st, v_st)


def test_config2_shaped_step_vs_oracle(cuda):
    """One full step in the shape of BASELINE.json configs[1] - G = 16 blocks + suffix (x4), D at 96x96 with
    its eight convs and the 18 432-wide head, MaskedVGG54 - at B=4 against the fp32 oracle's train_step
    (train.py:33-122): the three losses within 2 %, the fake batch >= 50 dB."""
    import sisr_b200 as m
    seed, lr = 910, 1e-5
    tr, (g_st, d_st, v_st) = _trainer(m, seed, 16, 0b10000, lr)
    hr = S.synthetic_hr(seed + 5, 4, 96)
    lr_img = O.lr_from_hr(hr, (24, 24))
    out = tr.step(hr.cuda(), lr_img.cuda())
    og, od = O.AdamState(O.trainable_names(g_st), lr), O.AdamState(O.trainable_names(d_st), lr)
    ref = O.train_step(g_st, d_st, v_st, hr, lr_img, d_strides=STRIDES, vgg_mask=0b10000, opt_g=og, opt_d=od)
    psnr = O.psnr(out["fake"].float().cpu(), ref["fake"])
    print(f"config-2 shape, B=4: PSNR {psnr:.1f} dB; " +
          "; ".join(f"{k} {float(out[k]):.5f} (oracle {ref[k]:.5f})" for k in ("err_d", "err_g_adv", "err_g_cont")))
    assert psnr >= 50.0, psnr
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(out[k]) - ref[k]) < 2e-2 * abs(ref[k]), (k, float(out[k]), ref[k])
    # second step on the updated weights (lr 1e-5: the update is far below bf16 resolution of the losses)
    hr2 = S.synthetic_hr(seed + 6, 4, 96)
    lr2 = O.lr_from_hr(hr2, (24, 24))
    out2 = tr.step(hr2.cuda(), lr2.cuda())
    ref2 = O.train_step(g_st, d_st, v_st, hr2, lr2, d_strides=STRIDES, vgg_mask=0b10000, opt_g=og, opt_d=od)
    print("step 2: " + "; ".join(f"{k} {float(out2[k]):.5f} (oracle {ref2[k]:.5f})"
                                 for k in ("err_d", "err_g_adv", "err_g_cont")))
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(out2[k]) - ref2[k]) < 2e-2 * abs(ref2[k]), (k, float(out2[k]), ref2[k])


def test_frozen_trunk_step_vs_oracle(cuda):
    """BASELINE.json configs[3]: x2 weights wrapped by GeneratorSuffix(freeze_prefix=True, freeze_upscale=True,
    freeze_end=True) (config.py:96, model_generator.py:103-115, 130-131): losses, the three trainable
    tensors' gradients and their update against O.train_step(g_trainable=...); the frozen trunk still
    advances its BN running statistics and spectral-norm vectors."""
    import sisr_b200 as m
    seed, lr = 920, 1e-3
    tr, (g_st, d_st, v_st) = _trainer(m, seed, 16, 0b10000, lr, frozen=True)
    trainable = sorted(k for k, p in tr.net_g.named_parameters() if p.requires_grad)
    assert trainable == ["upscale.0.bias", "upscale.0.weight_orig", "upscale.2.weight"]
    assert sum(p.numel() for p in tr.net_g.parameters() if p.requires_grad) == 147713
    g_init = S.clone_state(g_st)
    hr = S.synthetic_hr(seed + 5, 4, 96)
    lr_img = O.lr_from_hr(hr, (24, 24))
    out = tr.step(hr.cuda(), lr_img.cuda())

    def oracle(emulate, jitter=None):
        gs = S.clone_state(g_init)
        ds = S.discriminator_state(seed + 1, (3, 96, 96), FEATS, STRIDES)
        with contextlib.ExitStack() as es:
            if emulate:
                es.enter_context(O.emulate_bf16_storage())
            if jitter is not None:
                es.enter_context(O.jitter_before_rounding(1e-6, jitter))
            r = O.train_step(gs, ds, S.vgg_state(seed + 2, 0b10000), hr, lr_img, d_strides=STRIDES,
                             vgg_mask=0b10000, opt_g=O.AdamState(trainable, lr),
                             opt_d=O.AdamState(O.trainable_names(ds), lr), g_trainable=trainable)
        r["g_after"] = gs
        return r
    ref32, ref, j1, j2 = oracle(False), oracle(True), oracle(True, 1), oracle(True, 2)
    assert O.psnr(out["fake"].float().cpu(), ref32["fake"]) >= 50.0
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        assert abs(float(out[k]) - ref32[k]) < 2e-2 * abs(ref32[k]), (k, float(out[k]), ref32[k])
    grads = {k: p.grad for k, p in tr.net_g.named_parameters()}
    for k in trainable:
        floor = O.rel_l2(j1["g_grads"][k], j2["g_grads"][k])
        err = rel(grads[k], ref["g_grads"][k])
        print(f"frozen trunk: grad {k} vs bf16-oracle {err:.4f} (floor {floor:.4f})")
        bound = max(1e-1, 2.5 * floor) if grads[k].numel() == 1 else max(8e-2, 1.5 * floor)    # scalar: see above
        assert err < bound, (k, err, floor)
    for k, p in tr.net_g.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, k
            assert torch.equal(p.detach().cpu(), g_init[k]), k          # frozen weights untouched
    sd = tr.net_g.state_dict()
    after = ref32["g_after"]
    for k in ("base.block_list.7.layers.1.running_mean", "base.block_list.15.layers.4.running_var",
              "base.block_list_end.1.running_mean"):
        assert not torch.equal(sd[k].cpu(), g_init[k]), k                # statistics still advance
        assert rel(sd[k], after[k]) < 2e-2, (k, rel(sd[k], after[k]))
    for k in ("base.block_list.3.layers.0.weight_u", "base.first_layers.0.weight_v"):
        assert rel(sd[k], after[k]) < 1e-4, k
    # the suffix conv moved the way the oracle's Adam moved it (sign agreement, first step is sign-like)
    k = "upscale.0.weight_orig"
    mine = sd[k].cpu() - g_init[k]
    want = ref["g_after"][k] - g_init[k]
    same = float((torch.sign(mine) == torch.sign(want)).float().mean())
    floor = float((torch.sign(j1["g_after"][k] - g_init[k]) == torch.sign(j2["g_after"][k] - g_init[k])).float().mean())
    assert same > min(0.9, floor - 0.05), (same, floor)
