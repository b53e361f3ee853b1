"""Fused multi-tensor Adam with a device-side step counter and learning-rate schedule.

Replaces ``optim.Adam(net.parameters(), lr, betas=(.9, .999))`` + ``LambdaLR(f ** iteration)``
(config.py:170-180, 293-294; train.py:75,108,121-122).  It is a ``torch.optim.Optimizer`` so host
schedulers still work in eager mode; with ``decay_per_step`` the schedule runs on the device, which
lets a captured CUDA graph of the whole training step replay with the right learning rate.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import call


# keys of a torch.optim.Adam parameter group that this optimizer does not use but carries, so that a
# state_dict written here loads into torch.optim.Adam and steps there (and the other way round)
_TORCH_ADAM_GROUP = dict(weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                         differentiable=False, fused=None, decoupled_weight_decay=False)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, decay_per_step=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, decay_per_step=decay_per_step,
                                      **_TORCH_ADAM_GROUP))
        self._dev_state = {}
        self._sched_origin = {}         # group index -> Adam step count at which the LR schedule (re)started
        self.grad_scale = 1.0           # set to 1/world_size by the data-parallel gradient sync
        self.grad_views = None          # optional {param: flat-bucket view} provided by GradSync

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("decay_per_step", self.defaults.get("decay_per_step", 1.0))
            for k, v in _TORCH_ADAM_GROUP.items():
                group.setdefault(k, v)

    def _device_scalars(self, gi, device):
        st = self._dev_state.get(gi)
        if st is None or st[0].device != device:
            st = (torch.zeros(1, dtype=torch.int32, device=device),
                  torch.zeros(4, dtype=torch.float32, device=device))
            self._dev_state[gi] = st
        return st

    # -- checkpoint interchange with torch.optim.Adam (utils.py:107-112, config.py:296-302) -----------
    def state_dict(self):
        """torch.optim.Adam layout: per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` (the step
        counter lives on the device here; it is copied into every parameter's state on export) and
        parameter groups with every key torch.optim.Adam reads.  ``lr`` is the schedule's initial
        value (the decay runs on the device and never rewrites the group), which is also what the
        reference resumes from: its ``LambdaLR`` is rebuilt at epoch 0 on resume (config.py:170-180)."""
        sd = super().state_dict()
        for gi, group in enumerate(sd["param_groups"]):
            st = self._dev_state.get(gi)
            step = int(st[0].item()) if st is not None else 0
            group.setdefault("initial_lr", group["lr"])
            for idx in group["params"]:
                if idx in sd["state"]:
                    sd["state"][idx]["step"] = torch.tensor(float(step))
        return sd

    def load_state_dict(self, state_dict):
        """Accepts a state written by this class or by ``torch.optim.Adam`` (+ ``LambdaLR``).  Saved
        groups are merged OVER the current ones (a torch group has no ``decay_per_step``; ours keeps
        the value the trainer configured).  The learning rate restarts from the schedule's initial
        value - ``initial_lr`` when a scheduler wrote it, else ``lr`` - exactly as the reference does
        when it builds a fresh ``LambdaLR`` after loading the optimizers (config.py:296-302, 340-343);
        Adam's own step count (bias correction) continues."""
        import copy
        state_dict = copy.deepcopy(state_dict)
        mine = self.param_groups
        for gi, saved in enumerate(state_dict["param_groups"]):
            cur = mine[gi] if gi < len(mine) else self.defaults
            for k, v in cur.items():
                if k != "params":
                    saved.setdefault(k, v)
            if "initial_lr" in saved:
                saved["lr"] = saved["initial_lr"]
        super().load_state_dict(state_dict)
        for gi, group in enumerate(self.param_groups):
            steps = [int(self.state[p]["step"]) for p in group["params"]
                     if p in self.state and "step" in self.state[p]]
            dev = next((p.device for p in group["params"] if p.is_cuda), None)
            if steps and dev is not None:
                step_t, _ = self._device_scalars(gi, dev)
                step_t.fill_(max(steps))
                self._sched_origin[gi] = max(steps)

    @torch.no_grad()
    def step(self, closure=None):
        for gi, group in enumerate(self.param_groups):
            ps, gs, ms, vs = [], [], [], []
            for p in group["params"]:
                g = self.grad_views[p] if (self.grad_views is not None and p in self.grad_views) else p.grad
                if g is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("sisr_b200.optim.Adam only runs on CUDA tensors")
                state = self.state[p]
                if not state:
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ps.append(p)
                gs.append(g.contiguous())
                ms.append(state["exp_avg"])
                vs.append(state["exp_avg_sq"])
            if not ps:
                continue
            step_t, hyper = self._device_scalars(gi, ps[0].device)
            stream = torch.cuda.current_stream().cuda_stream
            b1, b2 = group["betas"]
            decay = float(group["decay_per_step"])
            # the kernel computes lr0 * decay**(t - 1) from the Adam step count t; after a resume the
            # schedule counts from the step at which the state was loaded
            lr0 = float(group["lr"]) / decay ** self._sched_origin.get(gi, 0)
            call("sisr_adam_tick", step_t, lr0, decay, b1, b2, hyper, stream)
            n = len(ps)
            arr = ctypes.c_void_p * n
            numel = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
            call("sisr_adam_multi", n, arr(*[p.data_ptr() for p in ps]), arr(*[g.data_ptr() for g in gs]),
                 arr(*[m.data_ptr() for m in ms]), arr(*[v.data_ptr() for v in vs]), numel, hyper,
                 b1, b2, float(group["eps"]), float(self.grad_scale), stream)
            self._keepalive = gs
        return None
