"""Diagnostic (GPU): discriminator input-gradient error vs the bf16-storage-emulating oracle, next to
the distance between two jittered oracle runs (the noise floor), over several seeds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sisr_b200 as m
from sisr_b200 import ops
from oracle import srgan_oracle as O
from oracle import state_factory as S

def rel(a, b): return O.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())

shape, feats, strides = (3, 32, 32), [64, 64, 128, 128, 256, 256], [1, 2, 1, 2, 1, 2]
for seed in range(900, 906):
    st = S.discriminator_state(seed, shape, feats, strides)
    net = m.Discriminator(shape, feats, strides)
    torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    net = net.cuda().train()
    x0 = S.synthetic_hr(seed + 7, 4, 32)
    x = x0.cuda().requires_grad_(True)
    out = net(x)
    loss, _ = ops.bce_loss(out.view(-1), 0.9)
    loss.backward()
    grads = {k: p.grad for k, p in net.named_parameters()}

    def emulated(js=None):
        emu = S.discriminator_state(seed, shape, feats, strides)
        names = O.trainable_names(emu)
        leaf = O._leaf(emu, names)
        xe = x0.clone().requires_grad_(True)
        with O.emulate_bf16_storage():
            jit = O.jitter_before_rounding(1e-6, js) if js is not None else None
            if jit: jit.__enter__()
            oe = O.discriminator_forward(leaf, xe, strides, True)
            ge = torch.autograd.grad(O.bce(oe.view(-1), 0.9), [xe] + [leaf[k] for k in names])
            if jit: jit.__exit__()
        return names, oe, ge
    names, oe, ge = emulated()
    _, _, j1 = emulated(1)
    _, _, j2 = emulated(2)
    _, _, j3 = emulated(3)
    print(f"seed {seed}: out rel {rel(out, oe):.2e}  dx err {rel(x.grad, ge[0]):.4f}  "
          f"floor(j1,j2) {rel(j1[0], j2[0]):.4f} floor(j1,ref) {rel(j1[0], ge[0]):.4f} floor(j3,ref) {rel(j3[0], ge[0]):.4f}")
    i = names.index("fc.0.weight") + 1
    print(f"     fc.0.weight err {rel(grads['fc.0.weight'], ge[i]):.4f} floor {rel(j1[i], j2[i]):.4f}; "
          f"conv.0.weight_orig err {rel(grads['conv.0.weight_orig'], ge[1 + names.index('conv.0.weight_orig')]):.4f} "
          f"floor {rel(j1[1 + names.index('conv.0.weight_orig')], j2[1 + names.index('conv.0.weight_orig')]):.4f}")
