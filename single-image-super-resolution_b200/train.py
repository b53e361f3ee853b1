"""The SRGAN training step (reference: train.py:33-122, 128-186), same order of operations:

  fake = G(lr)
  D update : BCE(D(real), 0.9) + sum over [fake.detach(), replayed fakes] of BCE(D(.), 0) -> Adam
  G update : 5e-2 * BCE(D(fake), 1) + 1.0 * mean((E(real) - E(fake))**2)               -> Adam

Differences that do not change results: no ``.item()`` host syncs inside the step (losses are
returned as device scalars), the replay list lives on the GPU, and the D weight gradient of the
G-update pass - which ``net_d.zero_grad()`` discards before any use - is not computed.
The whole step can be captured into one CUDA graph (``SRGANTrainer.capture``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import ops
from .optim import Adam


def gen_losses(content_loss_on_lr: bool = False, n_g=(0, float("inf")), n_d=(0, float("inf")),
               n_content=None, n_identity=None):
    """The epoch-scheduled loss weights of config.gen_losses (config.py:124-166), same defaults: each
    weight is active for epochs in [start, stop) and 0 outside (a zero weight SKIPS its branch of the step,
    train.py:56,86,94,106).  ``loss_weight_cont(epoch)`` returns (weight, kind) with kind "features"
    (the MaskedVGG extractor), "identity" (plain pixel MSE, weight x10) or None; in ``content_loss_on_lr``
    mode the defaults are adversarial 5e-3 and identity x10 x10 on the re-downsampled fake."""
    inf = float("inf")
    if n_content is None:
        n_content = (0, 0) if content_loss_on_lr else (0, inf)
    if n_identity is None:
        n_identity = (0, inf) if content_loss_on_lr else (0, 0)

    def loss_weight_adv_g(i):
        if n_g[0] <= i < n_g[1]:
            return 5e-3 if content_loss_on_lr else 5e-2
        return 0

    def loss_weight_adv_d(i):
        return 1.0 if n_d[0] <= i < n_d[1] else 0

    def loss_weight_cont(i):
        cont = n_content[0] <= i < n_content[1]
        iden = n_identity[0] <= i < n_identity[1]
        assert not cont or not iden
        f = 10.0 if content_loss_on_lr else 1.0
        if cont:
            return 1.0 * f, "features"
        if iden:
            return 10.0 * f, "identity"
        return 0, None

    return loss_weight_adv_g, loss_weight_adv_d, loss_weight_cont


@dataclass
class StepConfig:
    """Knobs of config.py that shape the step (config.py:24,38,49-54,124-166,186-188).  The three loss
    weights are either constants or functions of the epoch as ``gen_losses`` returns them (then
    ``loss_weight_cont(epoch)`` yields (weight, "features" | "identity" | None))."""
    lr: float = 1e-5
    betas: tuple = (0.9, 0.999)
    lr_decay_per_step: float = 1.0
    loss_weight_adv_g: object = 5e-2
    loss_weight_adv_d: object = 1.0
    loss_weight_cont: object = 1.0
    content_loss_on_lr: bool = False    # "unsupervised" variant (train.py:41-50, 95-97)
    real_label: float = 1.0
    real_label_reduced: float = 0.9
    fake_label: float = 0.0
    dis_list_old_len: int = 1000
    dis_list_old_freq: int = 1
    dis_list_old_ratio: float = 0.01
    use_replay: bool = True
    async_weight_grads: bool = True     # weight-gradient kernels on a side stream (ops.async_weight_grads)
    overlap_real_features: bool = True  # MaskedVGG(real) on a side stream, concurrent with G and the D update
    overlap_d_real: bool = True         # D(real) forward (and, through autograd, its backward) on a side stream:
                                        # it needs nothing from G, so it runs under G's latency-bound forward
    overlap_fake_features: bool = True  # MaskedVGG(fake) forward / backward on a side stream, concurrent with the
                                        # D update and the D pass of the G update


class SRGANTrainer:
    def __init__(self, net_g, net_d, extractor, cfg: Optional[StepConfig] = None, grad_sync=None):
        self.net_g, self.net_d, self.extractor = net_g, net_d, extractor
        self.cfg = cfg or StepConfig()
        c = self.cfg
        self.opt_g = Adam([p for p in net_g.parameters()], lr=c.lr, betas=c.betas,
                          decay_per_step=c.lr_decay_per_step)
        self.opt_d = Adam([p for p in net_d.parameters()], lr=c.lr, betas=c.betas,
                          decay_per_step=c.lr_decay_per_step)
        self.dis_list_old: List[torch.Tensor] = []
        self.iteration = 0
        self.grad_sync = grad_sync      # parallel.GradSync or None
        if grad_sync is not None:
            # replicas start identical (weights, spectral-norm vectors, BN statistics) and both optimizers
            # reduce their gradients through the bucketed all-reduce
            from .parallel import broadcast_module
            broadcast_module(extractor)
            grad_sync.attach(self.opt_d, net_d)
            grad_sync.attach(self.opt_g, net_g)
        self._graph = None              # graph of the plain step (no replayed fakes, epoch-0 weights)
        self._graphs = {}               # (replayed fakes, weights) -> (CUDAGraph, outputs)
        self._static = None
        self._static_old = []           # static bf16 buffers the replayed fakes are copied into
        self._pool = None
        self._side = None
        self._zero = None
        # parameters of D receive gradients from two streams by design (overlap_d_real): autograd's advice about
        # the accumulation stream does not apply
        warn_off = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if warn_off is not None and self.cfg.overlap_d_real:
            warn_off(False)

    def weights(self, epoch: int = 0):
        """(lw_adv_d, lw_adv_g, lw_cont, extractor kind) for ``epoch`` (config.py:136-164)."""
        c = self.cfg

        def val(w):
            return w(epoch) if callable(w) else w
        cont = val(c.loss_weight_cont)
        if not isinstance(cont, tuple):
            cont = (cont, "features" if cont else None)
        return val(c.loss_weight_adv_d), val(c.loss_weight_adv_g), cont[0], cont[1]

    # -- losses (train.py:128-186) -----------------------------------------------------------
    def adversarial_loss_d(self, real, curr_fake, old_fakes, d_real=None):
        """``d_real``: D(real) already computed (on a side stream that the current one has waited for)."""
        c = self.cfg
        if d_real is None:
            d_real = self.net_d(real).view(-1)
        err, d_x = ops.bce_loss(d_real, c.real_label_reduced)
        d_g_z1 = 0
        for fk in [curr_fake, *old_fakes]:
            d_fake = self.net_d(fk).view(-1)
            e, m = ops.bce_loss(d_fake, c.fake_label)
            err = err + e
            d_g_z1 = d_g_z1 + m
        return d_g_z1, d_x, err

    def adversarial_loss_g(self, fake):
        with ops.no_param_grads():
            out = self.net_d(fake).view(-1)
        err, d_g_z2 = ops.bce_loss(out, self.cfg.real_label)
        return d_g_z2, err

    def content_loss_g(self, real, fake, feat_real=None, kind="features", feat_fake=None):
        """train.py:183-186; ``kind`` "identity" = model_content_extractor.identity(): pixel MSE."""
        if kind == "identity":
            return ops.mse_loss(real, fake)
        a = self.extractor(real) if feat_real is None else feat_real
        b = self.extractor(fake) if feat_fake is None else feat_fake
        return ops.mse_loss(a, b)

    def _side_stream(self, which: int = 0):
        """0: MaskedVGG(real), 1: D(real), 2: MaskedVGG(fake)."""
        if self._side is None:
            self._side = {}
        if which not in self._side:
            self._side[which] = torch.cuda.Stream()
        return self._side[which]

    # -- one iteration -----------------------------------------------------------------------
    def step(self, img_hr: torch.Tensor, img_lr: torch.Tensor, old_fakes=(), epoch: int = 0, img_hr2=None):
        """One training iteration; every rank of a data-parallel run must call it (SyncBN statistics are
        exchanged only inside, see ops.sync_bn_scope).  Arguments: see ``_step``."""
        with ops.sync_bn_scope():
            return self._step(img_hr, img_lr, old_fakes, epoch, img_hr2)

    def _step(self, img_hr, img_lr, old_fakes=(), epoch: int = 0, img_hr2=None):
        """img_hr: (B,3,H,H) fp32 in [-1,1]; img_lr: (B,3,H/s,H/s).  Returns device scalars.
        ``epoch`` selects the scheduled loss weights; a zero weight skips its branch exactly as
        train.py:56-78, 85-102, 106-108 do.  ``img_hr2``: in ``content_loss_on_lr`` mode the HR batch of the
        second dataset that the discriminator sees as "real" (train.py:41-50); the content loss is then
        taken between ``img_lr`` and the re-downsampled fake (train.py:95-97)."""
        from .utils import lr_from_hr
        c = self.cfg
        lw_d, lw_g, lw_c, kind = self.weights(epoch)
        ops.begin_step(img_hr.device, track_weight_uses=c.async_weight_grads)
        if self._zero is None or self._zero.device != img_hr.device:
            self._zero = torch.zeros((), dtype=torch.float32, device=img_hr.device)
        on_lr = c.content_loss_on_lr
        real_d = img_hr2 if (on_lr and img_hr2 is not None) else img_hr
        cont_real = img_lr if on_lr else img_hr
        # the content-loss features of the REAL batch depend on nothing else in the step: they run on a
        # side stream, filling the SMs that the small generator / discriminator kernels leave idle
        feat_real = side = None
        if lw_c and kind == "features" and c.overlap_real_features and img_hr.is_cuda:
            side = self._side_stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                feat_real = self.extractor(cont_real)
        # D(real) needs nothing from G either.  Its forward runs on a second side stream under G's forward (a
        # latency-bound chain of small kernels); autograd then runs its backward on that stream too, next to the
        # backward of the D(fake) pass.  D(fake) starts only after D(real) has finished: the spectral-norm vectors
        # and BN running statistics advance in the reference's order (train.py:133-141).
        d_real = d_side = None
        # (not in data-parallel runs: with D's parameters fed from two streams the replicas of a CAPTURED step drifted
        # apart in the last bits - bench dp_check `weights_identical: false` at 2 ranks - although the eager step
        # passed tools/multi_gpu_check.py; cause not found within the round's GPU budget, profiles/r2_notes.md)
        single = self.grad_sync is None
        if lw_d and c.overlap_d_real and single and img_hr.is_cuda:
            self.net_d.zero_grad(set_to_none=True)
            d_side = self._side_stream(1)
            d_side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(d_side):
                d_real = self.net_d(real_d).view(-1)
        fake = self.net_g(img_lr)
        curr_fake = fake.detach()
        # ... and MaskedVGG(fake) does not depend on the D update: forward now, on a third side stream, concurrent
        # with the D update and the D pass of the G update; its backward (autograd: same stream) runs next to the
        # data-gradient pass through D
        feat_fake = f_side = None
        if lw_c and kind == "features" and c.overlap_fake_features and not on_lr and img_hr.is_cuda:
            f_side = self._side_stream(2)
            f_side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(f_side):
                feat_fake = self.extractor(fake)

        d_g_z1 = d_x = d_g_z2 = self._zero
        err_d = err_g_adv = err_g_cont = self._zero
        if lw_d:
            if d_side is not None:
                torch.cuda.current_stream().wait_stream(d_side)
            else:
                self.net_d.zero_grad(set_to_none=True)
            d_g_z1, d_x, err_d = self.adversarial_loss_d(real_d, curr_fake, old_fakes, d_real=d_real)
            err_d = err_d * lw_d
            with ops.async_weight_grads(c.async_weight_grads):
                err_d.backward()
            if d_side is not None:       # (the backward of the D(real) pass ran there)
                torch.cuda.current_stream().wait_stream(d_side)
            if self.grad_sync is not None:
                self.grad_sync.sync(self.opt_d)
            self.opt_d.step()

        self.net_g.zero_grad(set_to_none=True)
        if lw_g:
            d_g_z2, err_g_adv = self.adversarial_loss_g(fake)
            err_g_adv = err_g_adv * lw_g
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if f_side is not None:
            torch.cuda.current_stream().wait_stream(f_side)
        if lw_c:
            cont_fake = lr_from_hr(fake, tuple(img_lr.shape[-2:])) if on_lr else fake
            err_g_cont = self.content_loss_g(cont_real, cont_fake, feat_real, kind, feat_fake) * lw_c
        if lw_g or lw_c:
            with ops.async_weight_grads(c.async_weight_grads):
                (err_g_adv + err_g_cont).backward()
            if f_side is not None:       # (the backward of MaskedVGG(fake) ran there)
                torch.cuda.current_stream().wait_stream(f_side)
            if self.grad_sync is not None:
                self.grad_sync.sync(self.opt_g)
            self.opt_g.step()
        return {"fake": curr_fake, "err_d": err_d.detach(), "err_g_adv": err_g_adv.detach(),
                "err_g_cont": err_g_cont.detach(), "d_x": d_x, "d_g_z1": d_g_z1, "d_g_z2": d_g_z2}

    def _sample_old(self):
        import random
        c = self.cfg
        if not (c.use_replay and self.dis_list_old):
            return []
        k = int(len(self.dis_list_old) * c.dis_list_old_ratio)          # train.py:145
        return [self.dis_list_old[i] for i in random.sample(range(len(self.dis_list_old)), k)]

    def _remember(self, fake):
        import random
        c = self.cfg
        if c.use_replay and self.iteration % c.dis_list_old_freq == 0:   # train.py:66-71
            snap = fake.to(torch.bfloat16)      # replay list kept on the GPU in bf16 (a clone: ``fake`` may be a static graph output)
            if snap.data_ptr() == fake.data_ptr():
                snap = snap.clone()
            if len(self.dis_list_old) == c.dis_list_old_len:
                self.dis_list_old[random.randint(0, c.dis_list_old_len - 1)] = snap
            else:
                self.dis_list_old.append(snap)
        self.iteration += 1

    def train_iteration(self, img_hr, img_lr, epoch: int = 0, img_hr2=None, graph: bool = False):
        """step() plus the reference's experience-replay bookkeeping (train.py:59-71,144-146).  With
        ``graph=True`` the step is replayed from a CUDA graph captured for this number of replayed fakes
        and this epoch's loss weights (captured on first use, see ``capture``), so that training past
        iteration 100 - when int(len * 0.01) >= 1 old fake batches join every D update - stays on the
        graph path."""
        old = self._sample_old()
        if graph:
            out = self.replay(img_hr, img_lr, old_fakes=old, epoch=epoch, img_hr2=img_hr2)
        else:
            out = self.step(img_hr, img_lr, [o.float() for o in old], epoch=epoch, img_hr2=img_hr2)
        self._remember(out["fake"])
        return out

    # -- checkpoint (same dictionary as utils._save, utils.py:107-114) -----------------------------
    def checkpoint(self, epoch: int = 0) -> dict:
        """The replay list is exported the way the reference keeps it (fp32 CPU tensors,
        config.py:53, train.py:60-61), so that ``gen_dis_list`` + ``adversarial_loss_d`` of the reference
        can resume from this dictionary; in memory it stays bf16 on the GPU."""
        return {"epoch": epoch, "net_g": self.net_g.state_dict(), "net_d": self.net_d.state_dict(),
                "opti_g": self.opt_g.state_dict(), "opti_d": self.opt_d.state_dict(),
                "dis_list": [t.float().cpu() for t in self.dis_list_old]}

    def restore(self, checkpoint: dict) -> int:
        """Resume as config.py does (config.py:90-92, 296-302, 308-331): non-strict weight loading
        (``module.`` prefixes of a DataParallel checkpoint are stripped), optimizer state best
        effort, replay list as saved.  Returns the starting epoch."""
        def strip(sd):
            return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        if "net_g" in checkpoint:
            self.net_g.load_state_dict(strip(checkpoint["net_g"]), strict=False)
        if "net_d" in checkpoint:
            self.net_d.load_state_dict(strip(checkpoint["net_d"]), strict=False)
        for opt, key in ((self.opt_g, "opti_g"), (self.opt_d, "opti_d")):
            if key in checkpoint:
                try:
                    opt.load_state_dict(checkpoint[key])
                except Exception as e:  # noqa: BLE001 - mirrors the reference's best-effort load
                    print("erreur chargement optimizers:", e)
        dev = next(self.net_g.parameters()).device
        self.dis_list_old = [t.to(dev).to(torch.bfloat16) for t in checkpoint.get("dis_list", [])]
        self._graph = None          # a captured graph holds the old optimizer buffers
        self._graphs = {}
        return int(checkpoint.get("epoch", 0))

    # -- CUDA graph of the whole step ----------------------------------------------------------
    def _graph_key(self, n_old, epoch):
        return (n_old,) + tuple(self.weights(epoch))

    def capture(self, img_hr: torch.Tensor, img_lr: torch.Tensor, warmup: int = 2, n_old: int = 0,
                epoch: int = 0, img_hr2=None):
        """Capture ``step`` into one CUDA graph with static input buffers: one graph per (number of
        replayed fakes, loss weights of the epoch).  All graphs share one memory pool (they never run
        concurrently) and the static inputs.  ``warmup`` eager steps run first (they ARE training steps);
        the graphs for n_old > 0 are normally captured with warmup=0 by ``replay`` on first use."""
        if self._static is None:
            self._static = (img_hr.clone(), img_lr.clone(), img_hr.clone() if img_hr2 is None else img_hr2.clone())
        while len(self._static_old) < n_old:
            self._static_old.append(torch.zeros(self._static[0].shape, dtype=torch.bfloat16, device=img_hr.device))
        hr2 = self._static[2] if self.cfg.content_loss_on_lr else None

        def run():
            return self.step(self._static[0], self._static[1], [o.float() for o in self._static_old[:n_old]],
                             epoch=epoch, img_hr2=hr2)
        if warmup:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    run()
            torch.cuda.current_stream().wait_stream(side)
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self._pool):
            out = run()
        key = self._graph_key(n_old, epoch)
        self._graphs[key] = (graph, out)
        if key == self._graph_key(0, 0):
            self._graph, self._graph_out = graph, out
        return out

    def replay_from_feed(self, feed: "HostFeed"):
        """One captured step on the next batch of ``feed`` (pinned host HR patches): the LR input is
        made on the device (utils.lr_from_hr, train.py:46), the H2D copy of the following batch runs
        on the feed's copy stream meanwhile."""
        from .utils import lr_from_hr
        hr_dev, slot = feed.take()
        self._static[0].copy_(hr_dev, non_blocking=True)
        feed.release(slot)
        size = tuple(self._static[1].shape[-2:])
        self._static[1].copy_(lr_from_hr(self._static[0], size), non_blocking=True)
        self._graph.replay()
        return self._graph_out

    def replay(self, img_hr: torch.Tensor, img_lr: torch.Tensor, old_fakes=(), epoch: int = 0, img_hr2=None):
        """Replay the graph of this (number of replayed fakes, epoch weights); captured on first use."""
        key = self._graph_key(len(old_fakes), epoch)
        if key not in self._graphs:
            if self._static is None:
                raise RuntimeError("SRGANTrainer.replay: call capture() first")
            self.capture(self._static[0], self._static[1], warmup=0, n_old=len(old_fakes), epoch=epoch)
        graph, out = self._graphs[key]
        self._static[0].copy_(img_hr, non_blocking=True)
        self._static[1].copy_(img_lr, non_blocking=True)
        if img_hr2 is not None:
            self._static[2].copy_(img_hr2, non_blocking=True)
        for buf, o in zip(self._static_old, old_fakes):
            buf.copy_(o, non_blocking=True)
        graph.replay()
        return out


class HostFeed:
    """Double-buffered pinned-host -> device input pipeline for ``SRGANTrainer.replay_from_feed``:
    ``submit`` starts the asynchronous H2D copy of a batch on a copy stream, ``take`` hands the oldest
    submitted batch to the current stream.  With one batch submitted ahead, the copy of batch i+1
    overlaps the compute of step i (the reference's DataLoader + ``.to(device)``, train.py:40-46,
    is synchronous)."""

    def __init__(self, shape, device, depth: int = 2):
        self.stage = [torch.empty(shape, dtype=torch.float32, device=device) for _ in range(depth)]
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.stream = torch.cuda.Stream(device=device)
        self.submitted = self.taken = 0

    def submit(self, hr_host: torch.Tensor):
        if self.submitted - self.taken >= len(self.stage):
            raise RuntimeError("HostFeed: every stage buffer holds a batch that has not been taken")
        b = self.submitted % len(self.stage)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[b])       # the step that last read this buffer
            self.stage[b].copy_(hr_host, non_blocking=True)
            self.copied[b].record(self.stream)
        self.submitted += 1

    def take(self):
        if self.taken >= self.submitted:
            raise RuntimeError("HostFeed: nothing submitted")
        b = self.taken % len(self.stage)
        torch.cuda.current_stream().wait_event(self.copied[b])
        self.taken += 1
        return self.stage[b], b

    def release(self, slot: int):
        self.free[slot].record(torch.cuda.current_stream())
