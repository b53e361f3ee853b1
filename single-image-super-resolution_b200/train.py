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


@dataclass
class StepConfig:
    """Knobs of config.py that shape the step (config.py:38,49-54,136-164,186-188)."""
    lr: float = 1e-5
    betas: tuple = (0.9, 0.999)
    lr_decay_per_step: float = 1.0
    loss_weight_adv_g: float = 5e-2
    loss_weight_adv_d: float = 1.0
    loss_weight_cont: float = 1.0
    real_label: float = 1.0
    real_label_reduced: float = 0.9
    fake_label: float = 0.0
    dis_list_old_len: int = 1000
    dis_list_old_freq: int = 1
    dis_list_old_ratio: float = 0.01
    use_replay: bool = True
    async_weight_grads: bool = True     # weight-gradient kernels on a side stream (ops.async_weight_grads)
    overlap_real_features: bool = True  # MaskedVGG(real) on a side stream, concurrent with G and the D update


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
        self._graph = None
        self._static = None
        self._side = None

    # -- losses (train.py:128-186) -----------------------------------------------------------
    def adversarial_loss_d(self, real, curr_fake, old_fakes):
        c = self.cfg
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

    def content_loss_g(self, real, fake, feat_real=None):
        a = self.extractor(real) if feat_real is None else feat_real
        b = self.extractor(fake)
        return ops.mse_loss(a, b)

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    # -- one iteration -----------------------------------------------------------------------
    def step(self, img_hr: torch.Tensor, img_lr: torch.Tensor, old_fakes=()):
        """img_hr: (B,3,H,H) fp32 in [-1,1]; img_lr: (B,3,H/s,H/s).  Returns device scalars."""
        c = self.cfg
        ops.begin_step(img_hr.device, track_weight_uses=c.async_weight_grads)
        # the content-loss features of the REAL batch depend on nothing else in the step: they run on a
        # side stream, filling the SMs that the small generator / discriminator kernels leave idle
        feat_real = side = None
        if c.overlap_real_features and img_hr.is_cuda:
            side = self._side_stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                feat_real = self.extractor(img_hr)
        fake = self.net_g(img_lr)

        self.net_d.zero_grad(set_to_none=True)
        curr_fake = fake.detach()
        d_g_z1, d_x, err_d = self.adversarial_loss_d(img_hr, curr_fake, old_fakes)
        err_d = err_d * c.loss_weight_adv_d
        with ops.async_weight_grads(c.async_weight_grads):
            err_d.backward()
        if self.grad_sync is not None:
            self.grad_sync.sync(self.opt_d)
        self.opt_d.step()

        self.net_g.zero_grad(set_to_none=True)
        d_g_z2, err_g_adv = self.adversarial_loss_g(fake)
        err_g_adv = err_g_adv * c.loss_weight_adv_g
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        err_g_cont = self.content_loss_g(img_hr, fake, feat_real) * c.loss_weight_cont
        with ops.async_weight_grads(c.async_weight_grads):
            (err_g_adv + err_g_cont).backward()
        if self.grad_sync is not None:
            self.grad_sync.sync(self.opt_g)
        self.opt_g.step()
        return {"fake": curr_fake, "err_d": err_d.detach(), "err_g_adv": err_g_adv.detach(),
                "err_g_cont": err_g_cont.detach(), "d_x": d_x, "d_g_z1": d_g_z1, "d_g_z2": d_g_z2}

    def train_iteration(self, img_hr, img_lr):
        """step() plus the reference's experience replay bookkeeping (train.py:59-71,144-146)."""
        import random
        c = self.cfg
        old = []
        if c.use_replay and self.dis_list_old:
            k = int(len(self.dis_list_old) * c.dis_list_old_ratio)
            old = [self.dis_list_old[i].float() for i in random.sample(range(len(self.dis_list_old)), k)]
        out = self.step(img_hr, img_lr, old)
        if c.use_replay and self.iteration % c.dis_list_old_freq == 0:
            snap = out["fake"].to(torch.bfloat16)      # replay list kept on the GPU in bf16
            if len(self.dis_list_old) == c.dis_list_old_len:
                self.dis_list_old[random.randint(0, c.dis_list_old_len - 1)] = snap
            else:
                self.dis_list_old.append(snap)
        self.iteration += 1
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
        return int(checkpoint.get("epoch", 0))

    # -- CUDA graph of the whole step ----------------------------------------------------------
    def capture(self, img_hr: torch.Tensor, img_lr: torch.Tensor, warmup: int = 2):
        """Capture ``step`` (no replayed fakes) into one CUDA graph with static input buffers."""
        self._static = (img_hr.clone(), img_lr.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_out = self.step(*self._static)
        return self._graph_out

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

    def replay(self, img_hr: torch.Tensor, img_lr: torch.Tensor):
        self._static[0].copy_(img_hr, non_blocking=True)
        self._static[1].copy_(img_lr, non_blocking=True)
        self._graph.replay()
        return self._graph_out


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
