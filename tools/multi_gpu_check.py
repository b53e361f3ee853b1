"""Multi-GPU parity check (run with torchrun, one rank per GPU):
  1. the NVLink peer-memory all-reduce equals the sum of the ranks' vectors, bit-identical on all ranks;
  2. one data-parallel SRGAN step (batch sharded, SyncBN over peer memory, bucketed NCCL gradient
     all-reduce) gives the same losses and post-step weights as ONE process on the whole batch.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import sisr_b200 as m
from sisr_b200 import ops, parallel
from oracle import state_factory as S
import torch.nn.functional as F


def build(dev, grad_sync, seed=800, shape=(3, 32, 32), feats=(64, 64, 128, 128), strides=(1, 2, 1, 2), mask=0b00010):
    g_st = S.generator_state(seed, n_blocks=2, n_suffix=1)
    d_st = S.discriminator_state(seed + 1, shape, list(feats), list(strides))
    v_st = S.vgg_state(seed + 2, mask)
    net_g = m.GeneratorSuffix(m.Generator(2, 64, 256, [2], use_sn=True))
    net_d = m.Discriminator(shape, list(feats), list(strides))
    ext = m.MaskedVGG(mask)
    for net, st in ((net_g, g_st), (net_d, d_st), (ext, v_st)):
        torch.nn.Module.load_state_dict(net, S.clone_state(st), strict=True)
    net_g, net_d, ext = net_g.to(dev), net_d.to(dev), ext.to(dev)
    cfg = m.StepConfig(lr=1e-5, use_replay=False,
                       async_weight_grads=os.environ.get("MGC_ASYNC", "1") != "0",          # diagnostics
                       overlap_real_features=os.environ.get("MGC_OVERLAP", "1") != "0")
    return m.SRGANTrainer(net_g, net_d, ext, cfg, grad_sync=grad_sync)


def main():
    rank, local, world = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ok = True
    # 1. peer all-reduce
    px = ops._peer
    assert px is not None and world > 1
    for n in (1, 129, 1025):
        v = torch.arange(n, device=dev, dtype=torch.float32) * (rank + 1) + 0.25 * rank
        px.reset()
        for _ in range(3):          # same slot reused across "steps": epochs must advance
            w = v.clone()
            ops.call("sisr_peer_allreduce", px.bases, px.rank, px.world, 7, w, n, ops._stream())
            torch.cuda.synchronize()
            dist.barrier()
        want = sum(torch.arange(n, dtype=torch.float32) * (r + 1) + 0.25 * r for r in range(world))
        good = torch.equal(w.cpu(), want)
        ok &= good
        if rank == 0:
            print(f"peer_allreduce n={n}: {'ok' if good else 'MISMATCH'}", flush=True)
    # 2. data-parallel step vs single process on the global batch
    per = 4
    hr_all = S.synthetic_hr(4321, per * world, 32)
    lr_all = F.interpolate(hr_all, (8, 8), mode="bicubic", align_corners=True).clamp(-1, 1)
    # 1b. SyncBN forward + backward in isolation (no NCCL, no side streams): a chain of BatchNorm layers on
    # per-rank shards against one process on the concatenated batch
    from sisr_b200.ops import BnActFn, BnCfg, ACT_PRELU
    gen = torch.Generator().manual_seed(99)
    xs = (torch.randn(world * 4, 8, 8, 64, generator=gen) * 1.5 + 0.3).to(torch.bfloat16)
    gs = torch.randn(world * 4, 8, 8, 64, generator=gen).to(torch.bfloat16)

    def bn_chain(x, g, sync):
        x = x.to(dev).requires_grad_(True)
        gam = torch.full((64,), 1.2, device=dev, requires_grad=True)
        bet = torch.full((64,), 0.1, device=dev, requires_grad=True)
        slope = torch.full((1,), 0.25, device=dev, requires_grad=True)
        h = x
        ops.begin_step(dev)
        with (ops.sync_bn_scope() if sync else __import__("contextlib").nullcontext()):
            for _ in range(6):
                rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
                h = BnActFn.apply(h, None, gam, bet, rm, rv, torch.zeros((), dtype=torch.long, device=dev), None, slope,
                                  BnCfg(act=ACT_PRELU, training=True))
            h.backward(g.to(dev))
        torch.cuda.synchronize()
        return h.detach().float(), x.grad.float(), gam.grad.clone()
    sl4 = slice(rank * 4, (rank + 1) * 4)
    out_dp, dx_dp, dgam_dp = bn_chain(xs[sl4], gs[sl4], True)
    dist.all_reduce(dgam_dp)
    if rank == 0:
        keep = (ops._dist_group, ops._peer)
        ops.set_sync_group(None)
        ops.set_peer_exchange(None)
        out_1p, dx_1p, dgam_1p = bn_chain(xs, gs, False)
        ops.set_sync_group(keep[0])
        ops.set_peer_exchange(keep[1])
        e_out = float((out_dp - out_1p[sl4]).norm() / out_1p[sl4].norm())
        e_dx = float((dx_dp - dx_1p[sl4]).norm() / dx_1p[sl4].norm())
        e_g = float((dgam_dp - dgam_1p).norm() / dgam_1p.norm())
        print(f"isolated SyncBN chain (6 layers): out {e_out:.2e}  dx {e_dx:.2e}  dgamma {e_g:.2e}", flush=True)
        ok &= e_out < 2e-2 and e_dx < 2e-2 and e_g < 2e-2
    dist.barrier()
    tr = build(dev, parallel.GradSync(bucket_bytes=int(os.environ.get("MGC_BUCKET_MB", "1")) << 20))
    sl = slice(rank * per, (rank + 1) * per)
    def flat_grads(t, dp):
        """Flattened gradients the optimizers consumed in the last step (dp: all-reduced bucket views / world);
        also per parameter, for the diagnostics below."""
        out = {}
        for name, net, opt in (("G", t.net_g, t.opt_g), ("D", t.net_d, t.opt_d)):
            gs = []
            for pname, q in net.named_parameters():
                if not q.requires_grad:
                    continue
                if dp:
                    g_ = (opt.grad_views[q] / world).flatten()
                elif q.grad is not None:
                    g_ = q.grad.flatten()
                else:
                    g_ = torch.zeros(q.numel(), device=dev)
                gs.append(g_)
                out[name + ":" + pname] = g_.double().clone()
            out[name] = torch.cat(gs).double()
        return out

    outs = [tr.step(hr_all[sl].to(dev), lr_all[sl].to(dev))]
    torch.cuda.synchronize()
    g_dp = flat_grads(tr, True)             # gradients of step 1 (same weights on both sides)
    outs.append(tr.step(hr_all[sl].to(dev), lr_all[sl].to(dev)))
    torch.cuda.synchronize()
    losses = torch.stack([torch.stack([o["err_d"].reshape(()), o["err_g_adv"].reshape(()), o["err_g_cont"].reshape(())])
                          for o in outs])
    # the D / G adversarial losses are per-rank means over the local shard: average them
    dist.all_reduce(losses)
    losses /= world
    if rank == 0:
        ops.set_sync_group(None)
        ops.set_peer_exchange(None)
        ref = build(dev, None)
        routs = [ref.step(hr_all.to(dev), lr_all.to(dev))]
        torch.cuda.synchronize()
        g_1p = flat_grads(ref, False)
        routs.append(ref.step(hr_all.to(dev), lr_all.to(dev)))
        rl = torch.stack([torch.stack([o["err_d"].reshape(()), o["err_g_adv"].reshape(()), o["err_g_cont"].reshape(())])
                          for o in routs])
        rel = ((losses - rl).abs() / rl.abs())
        print("losses dp:", losses.tolist(), "\nlosses 1p:", rl.tolist(),
              f"\nmax rel diff step 1: {rel[0].max().item():.3e}  step 2: {rel[1].max().item():.3e}", flush=True)
        # lr = 1e-5 (config.py:38): Adam's first update is lr * sign(g) per element, so near-zero gradients
        # whose sign depends on the fp32 summation order move weights by 2 lr; with lr = 1e-3 that alone
        # made the G losses (evaluated AFTER the D update inside the same step) differ by 0.1-4 % from run
        # to run.  The synchronised gradients of step 1 are compared directly instead.
        ok &= rel[0].max().item() < 5e-3
        ok &= rel[1].max().item() < 2e-2
        for name in ("G", "D"):
            a, b = g_dp[name], g_1p[name]
            cosine = float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))
            ratio = float(a.norm() / b.norm().clamp_min(1e-30))
            print(f"step-1 gradient {name}: cosine {cosine:.5f}  norm ratio {ratio:.4f}", flush=True)
            # which parameters carry the difference: share of the squared difference per parameter
            diffs = []
            tot = float((a - b).pow(2).sum().clamp_min(1e-300))
            for k in g_dp:
                if k.startswith(name + ":"):
                    da, db = g_dp[k], g_1p[k]
                    diffs.append((float((da - db).pow(2).sum()) / tot, k, float(da.norm()), float(db.norm()), da.numel()))
            for share, k, na, nb, numel in sorted(diffs, reverse=True)[:6]:
                print(f"    {share * 100:5.1f} % of |dp - 1p|^2 in {k} (numel {numel}): |dp| {na:.4g}  |1p| {nb:.4g}", flush=True)
            vec = [k for k in g_dp if k.startswith(name + ":") and g_dp[k].numel() > 1]
            av, bv = torch.cat([g_dp[k] for k in vec]), torch.cat([g_1p[k] for k in vec])
            cos_v = float(av @ bv / (av.norm() * bv.norm()).clamp_min(1e-30))
            print(f"    without the scalar (PReLU slope) gradients: cosine {cos_v:.5f}  norm ratio "
                  f"{float(av.norm() / bv.norm().clamp_min(1e-30)):.4f}", flush=True)
            ok &= cosine > 0.98 and abs(ratio - 1.0) < 0.05
        worst = 0.0
        for (k, a), (_, b) in zip(tr.net_g.state_dict().items(), ref.net_g.state_dict().items()):
            if a.dtype.is_floating_point and "running" in k:
                worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)))
        print(f"BN running statistics (G) worst rel diff vs single process: {worst:.3e}", flush=True)
        ok &= worst < 4e-2
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() == 1.0 else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
