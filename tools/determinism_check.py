"""Run-to-run reproducibility of one training step (GPU): the same trainer is built TRIALS times from the same
seeds and stepped twice on the same inputs; per trial the losses, the gradients the optimizers consumed and the
post-step weights are compared with trial 0.  fp32 reductions whose order changes between runs (L2 `red.add`
of the small weight gradients, shared-memory atomics of the BatchNorm backward sums) move a gradient by a few
ulp; anything beyond that - a whole tensor off by percents - is a race, and this tool is how to find it.

usage: python tools/determinism_check.py [trials] [eager|graph]      (SISR_WGRAD_ATOMIC=0: split-K partial copies)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sisr_b200 as m
from oracle import srgan_oracle as O
from oracle import state_factory as S


def build(seed=700, shape=(3, 32, 32), feats=(64, 64, 128, 128), strides=(1, 2, 1, 2), mask=0b00010, lr=1e-3):
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
    d_st = S.discriminator_state(seed + 1, shape, list(feats), list(strides))
    v_st = S.vgg_state(seed + 2, mask)
    net_g = m.GeneratorSuffix(m.Generator(2, 64, 256, [2], use_sn=True))
    net_d = m.Discriminator(shape, list(feats), list(strides))
    ext = m.MaskedVGG(mask)
    for net, st in ((net_g, g_st), (net_d, d_st), (ext, v_st)):
        torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    return m.SRGANTrainer(net_g.cuda(), net_d.cuda(), ext.cuda(), m.StepConfig(lr=lr, use_replay=False))


def snapshot(tr):
    out = {}
    for name, net in (("G", tr.net_g), ("D", tr.net_d)):
        for k, p in net.named_parameters():
            if p.grad is not None:
                out[f"grad {name}:{k}"] = p.grad.detach().double().cpu().clone()
            out[f"weight {name}:{k}"] = p.detach().double().cpu().clone()
    return out


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
    hrs = [S.synthetic_hr(750 + i, 4, 32).cuda() for i in range(2)]
    lrs = [m.lr_from_hr(h, (8, 8)) for h in hrs]
    base = None
    worst_all = 0.0
    for t in range(trials):
        tr = build()
        snaps, losses = [], []
        if graph:
            tr.capture(hrs[0], lrs[0], warmup=1)
        for h, l in zip(hrs, lrs):
            o = tr.replay(h, l) if graph else tr.step(h, l)
            torch.cuda.synchronize()
            losses.append([float(o[k]) for k in ("err_d", "err_g_adv", "err_g_cont")])
            snaps.append(snapshot(tr))
        if base is None:
            base = (snaps, losses)
            print(f"trial 0 losses {losses}", flush=True)
            continue
        for s in range(2):
            bad = []
            for k, v in snaps[s].items():
                ref = base[0][s].get(k)
                if ref is None or ref.shape != v.shape:
                    continue
                den = float(ref.norm())
                e = float((v - ref).norm()) / den if den > 0 else float((v - ref).norm())
                if k.startswith("grad"):
                    worst_all = max(worst_all, e)
                if e > 1e-4:
                    bad.append((e, k, den))
            lo = [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(losses[s], base[1][s])]
            print(f"trial {t} step {s}: loss rel diff {['%.2e' % x for x in lo]}; tensors off by > 1e-4: {len(bad)}", flush=True)
            for e, k, den in sorted(bad, reverse=True)[:8]:
                print(f"      {e:9.3e}  {k}  (|ref| {den:.3e})", flush=True)
    print(f"worst gradient rel-L2 difference vs trial 0: {worst_all:.3e}", flush=True)


if __name__ == "__main__":
    main()
