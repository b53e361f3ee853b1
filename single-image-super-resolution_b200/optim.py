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


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, decay_per_step=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, decay_per_step=decay_per_step))
        self._dev_state = {}
        self.grad_scale = 1.0           # set to 1/world_size by the data-parallel gradient sync
        self.grad_views = None          # optional {param: flat-bucket view} provided by GradSync

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
        counter lives on the device here; it is copied into every parameter's state on export)."""
        sd = super().state_dict()
        for gi, group in enumerate(sd["param_groups"]):
            st = self._dev_state.get(gi)
            step = int(st[0].item()) if st is not None else 0
            for idx in group["params"]:
                if idx in sd["state"]:
                    sd["state"][idx]["step"] = torch.tensor(float(step))
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for gi, group in enumerate(self.param_groups):
            steps = [int(self.state[p]["step"]) for p in group["params"]
                     if p in self.state and "step" in self.state[p]]
            dev = next((p.device for p in group["params"] if p.is_cuda), None)
            if steps and dev is not None:
                step_t, _ = self._device_scalars(gi, dev)
                step_t.fill_(max(steps))

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
            call("sisr_adam_tick", step_t, float(group["lr"]), float(group["decay_per_step"]), b1, b2,
                 hyper, stream)
            n = len(ps)
            arr = ctypes.c_void_p * n
            numel = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
            call("sisr_adam_multi", n, arr(*[p.data_ptr() for p in ps]), arr(*[g.data_ptr() for g in gs]),
                 arr(*[m.data_ptr() for m in ms]), arr(*[v.data_ptr() for v in vs]), numel, hyper,
                 b1, b2, float(group["eps"]), float(self.grad_scale), stream)
            self._keepalive = gs
        return None
