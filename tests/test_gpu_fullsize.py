"""Parity properties at the FULL size of BASELINE.json configs[1] (batch 64, 96x96 HR), where the CPU
oracle would take minutes: size-independent properties of the kernels and an independent full-size
checker (PyTorch's own GPU convolution in fp32 - test infrastructure only)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
B = 64


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("shape", [(24, 64, 64, 1), (96, 64, 64, 1), (48, 128, 128, 1), (12, 512, 512, 1),
                                    (96, 64, 64, 2), (24, 64, 256, 1)],
                         ids=lambda s: "x".join(map(str, s)))
def test_conv_full_size_vs_torch_gpu_and_linearity(cuda, shape):
    """fprop / dgrad / wgrad of the step's conv shapes at batch 64 against torch's fp32 GPU conv on the
    same bf16-rounded operands (rel-L2 <= 1e-2), plus linearity conv(2*x1 - x2) = 2*conv(x1) - conv(x2)."""
    from sisr_b200 import ops
    h, cin, cout, stride = shape
    g = torch.Generator(device="cuda").manual_seed(h * 1000 + cin + cout + stride)
    x = bf(torch.randn(B, cin, h, h, device="cuda", generator=g))
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    ps = 2 if cout == 256 and cin == 64 else 0
    cfg = ops.ConvCfg(stride=stride, pad=1, ps_r=ps, want_stats=(ps == 0))

    def run(xin, need_grad=False):
        xd = xin.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).requires_grad_(need_grad)
        wd, bd = w.clone().requires_grad_(need_grad), b.clone().requires_grad_(need_grad)
        y, st = ops.Conv2dFn.apply(xd, wd, bd, None, None, None, cfg)
        return xd, wd, bd, y, st

    xd, wd, bd, y, st = run(x, True)
    xr, wr, br = x.clone().requires_grad_(True), bf(w).requires_grad_(True), b.clone().requires_grad_(True)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    y_ref = F.conv2d(xr, wr, br, stride=stride, padding=1)
    if ps:
        y_ref = F.pixel_shuffle(y_ref, 2)
    gy = bf(torch.randn(y_ref.shape, device="cuda", generator=g))
    y_ref.backward(gy)
    torch.backends.cudnn.allow_tf32 = prev
    assert rel(y.float().permute(0, 3, 1, 2), y_ref) < 1e-2
    if st is not None:                      # fused BN statistics = sums of the stored output
        yy = y.double()
        assert rel(st.double().sum(0)[:cout], yy.sum(dim=(0, 1, 2))) < 1e-3 or float(yy.sum().abs()) < 1
        assert rel(st.double().sum(0)[cout:], (yy * yy).sum(dim=(0, 1, 2))) < 1e-3
    y.backward(gy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    assert rel(xd.grad.float().permute(0, 3, 1, 2), xr.grad) < 1e-2
    assert rel(wd.grad, wr.grad) < 1e-2
    assert rel(bd.grad, br.grad) < 1e-2
    # linearity (bias removed): exact up to the bf16 rounding of inputs and outputs
    x2 = bf(torch.randn(B, cin, h, h, device="cuda", generator=g))
    b.zero_()
    ya = run(bf(2 * x - x2))[3].float()
    yb = 2 * run(x)[3].float() - run(x2)[3].float()
    assert rel(ya, yb) < 1.5e-2


def test_full_step_is_reproducible_and_bounded(cuda):
    """Two trainers built from the same seed give the same losses at batch 64 / 96x96 (only fp32
    atomics reorder), SR images stay in [-1, 1], BN running statistics move, SN vectors stay unit."""
    import bench
    from oracle import state_factory as S
    dev = torch.device("cuda")
    hr = S.synthetic_hr(99, B, 96).to(dev)
    lr = F.interpolate(hr, (24, 24), mode="bicubic", align_corners=True).clamp(-1, 1)
    outs = []
    for _ in range(2):
        tr = bench.build_trainer(dev, B, 1)
        o = tr.step(hr, lr)
        outs.append(o)
    for k in ("err_d", "err_g_adv", "err_g_cont"):
        a, b = float(outs[0][k]), float(outs[1][k])
        assert math.isfinite(a) and abs(a - b) < 2e-3 * abs(a), (k, a, b)
    fake = outs[0]["fake"]
    assert fake.shape == (B, 3, 96, 96) and float(fake.abs().max()) <= 1.0
    sd = tr.net_g.state_dict()
    assert float(sd["base.block_list.0.layers.1.running_mean"].abs().max()) > 0
    assert int(sd["base.block_list.0.layers.1.num_batches_tracked"]) == 1
    for k in ("base.block_list.3.layers.0.weight_u", "upscale.0.weight_v"):
        assert abs(float(sd[k].norm()) - 1.0) < 1e-4, k
