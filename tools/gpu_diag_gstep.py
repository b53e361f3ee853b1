"""Diagnostic (GPU): gradient of the G-step losses w.r.t. the fake image, per path, vs the oracle
(fp32 and bf16-storage-emulated), without any optimizer update."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sisr_b200 as m
from sisr_b200 import ops
from oracle import srgan_oracle as O
from oracle import state_factory as S

def rel(a, b): return O.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())
def cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))

seed, shape, feats, strides, mask = 700, (3, 32, 32), [64, 64, 128, 128, 256, 256], [1, 2, 1, 2, 1, 2], 0b00110
g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
d_st = S.discriminator_state(seed + 1, shape, feats, strides)
v_st = S.vgg_state(seed + 2, mask)
net_g = m.GeneratorSuffix(m.Generator(2, 64, 256, [2], use_sn=True))
net_d = m.Discriminator(shape, feats, strides)
ext = m.MaskedVGG(mask)
for net, st in ((net_g, g_st), (net_d, d_st), (ext, v_st)):
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
net_g, net_d, ext = net_g.cuda(), net_d.cuda(), ext.cuda()
hr = S.synthetic_hr(seed + 5, 4, 32); lr_img = O.lr_from_hr(hr, (8, 8))

fake = net_g(lr_img.cuda())
with ops.no_param_grads():
    out = net_d(fake).view(-1)
e_adv, _ = ops.bce_loss(out, 1.0)
(g_adv,) = torch.autograd.grad(e_adv, fake, retain_graph=True)
cont = ops.mse_loss(ext(hr.cuda()), ext(fake))
(g_cont,) = torch.autograd.grad(cont, fake, retain_graph=True)
names = O.trainable_names(g_st)
total = e_adv * 0.05 + cont
gg = dict(zip(names, torch.autograd.grad(total, [dict(net_g.named_parameters())[k] for k in names], allow_unused=True)))

for tag, emu in (("fp32", False), ("emu", True)):
    gs, ds, vs = S.clone_state(g_st), S.clone_state(d_st), S.clone_state(v_st)
    leaf = O._leaf(gs, names)
    ctx = O.emulate_bf16_storage() if emu else None
    if ctx: ctx.__enter__()
    f = O.generator_forward(leaf, lr_img, training=True)
    o = O.discriminator_forward(ds, f, strides, True).view(-1)
    ea = O.bce(o, 1.0)
    (ga,) = torch.autograd.grad(ea, f, retain_graph=True)
    c = torch.mean((O.masked_vgg_forward(vs, hr, mask) - O.masked_vgg_forward(vs, f, mask)) ** 2)
    (gc,) = torch.autograd.grad(c, f, retain_graph=True)
    gr = dict(zip(names, torch.autograd.grad(ea * 0.05 + c, [leaf[k] for k in names])))
    if ctx: ctx.__exit__()
    print(f"[{tag}] fake psnr {O.psnr(fake.detach().cpu(), f.detach()):.1f}  e_adv {float(e_adv):.5f}/{float(ea):.5f} cont {float(cont):.6f}/{float(c):.6f}")
    print(f"[{tag}] d(adv)/dfake rel {rel(g_adv, ga):.3f} cos {cos(g_adv, ga):.4f} |ref| {float(ga.norm()):.3e}")
    print(f"[{tag}] d(cont)/dfake rel {rel(g_cont, gc):.3f} cos {cos(g_cont, gc):.4f} |ref| {float(gc.norm()):.3e}")
    top = max(float(v.norm()) for v in gr.values())
    for k in names:
        if float(gr[k].norm()) > 1e-2 * top:
            print(f"[{tag}] G {k:46s} rel {rel(gg[k], gr[k]):.3f} cos {cos(gg[k], gr[k]):.4f}")
