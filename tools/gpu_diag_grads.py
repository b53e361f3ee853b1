"""Diagnostic (GPU): per-parameter gradient agreement of one training step against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sisr_b200 as m
from oracle import srgan_oracle as O
from oracle import state_factory as S

def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))

seed, shape, feats, strides, mask, lr = 700, (3, 32, 32), [64, 64, 128, 128, 256, 256], [1, 2, 1, 2, 1, 2], 0b00110, 1e-3
g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
d_st = S.discriminator_state(seed + 1, shape, feats, strides)
v_st = S.vgg_state(seed + 2, mask)
net_g = m.GeneratorSuffix(m.Generator(2, 64, 256, [2], use_sn=True))
net_d = m.Discriminator(shape, feats, strides)
ext = m.MaskedVGG(mask)
for net, st in ((net_g, g_st), (net_d, d_st), (ext, v_st)):
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
tr = m.SRGANTrainer(net_g.cuda(), net_d.cuda(), ext.cuda(), m.StepConfig(lr=lr, use_replay=False))
hr = S.synthetic_hr(seed + 5, 4, 32); lr_img = O.lr_from_hr(hr, (8, 8))
out = tr.step(hr.cuda(), lr_img.cuda())
ref = O.train_step(g_st, d_st, v_st, hr, lr_img, d_strides=strides, vgg_mask=mask,
                   opt_g=O.AdamState(O.trainable_names(g_st), lr), opt_d=O.AdamState(O.trainable_names(d_st), lr))
print("psnr fake", O.psnr(out["fake"].float().cpu(), ref["fake"]))
for k in ("err_d", "err_g_adv", "err_g_cont"):
    print(k, float(out[k]), ref[k])
for tag, net, rg in (("G", tr.net_g, ref["g_grads"]), ("D", tr.net_d, ref["d_grads"])):
    top = max(float(v.norm()) for v in rg.values())
    for k, p in net.named_parameters():
        r = rg[k]
        print(f"{tag} {k:48s} |ref|={float(r.norm()):.3e} ({float(r.norm())/top:.1e} of top) rel={O.rel_l2(p.grad.float().cpu(), r):.3f} cos={cos(p.grad, r):.4f}")
