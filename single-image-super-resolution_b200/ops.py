"""Autograd operators of the SRGAN step, each a thin wrapper that launches the CUDA kernels of
``libsisr_b200.so`` through the C ABI on the current stream.

Activations between operators are NHWC bf16 tensors; parameters, statistics and gradients of
parameters are fp32.  PyTorch supplies device memory, streams and the autograd tape only.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ConvDesc, call, query

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_PRELU, ACT_TANH = 0, 1, 2, 3, 4
LEAKY_SLOPE = 0.01   # nn.LeakyReLU() default used by the reference discriminator
SN_EPS = 1e-12
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

_skip_param_grads = False
# {data_ptr of a gradient tensor: (that tensor, its per-channel column sums)}: lets the kernel that
# writes a conv's output gradient also deliver the conv's bias gradient (cleared every step)
_colsum_cache = {}
# ReLU backward fused into the NEXT layer's dgrad (frozen VGG stack): a conv whose output went through
# a fused ReLU and that needs no parameter gradient registers its output here; the conv that consumes
# exactly that tensor masks its input gradient with it (sisr_conv_dgrad_masked) and lists the result
# in _premasked, so that the producer skips its own activation-backward pass.  ReLU only: the mask is
# idempotent, so a gradient that reaches the producer through a second path is still handled right.
_relu_outputs = {}
_premasked = {}


@contextlib.contextmanager
def no_param_grads():
    """Inside this context the operators do not compute parameter gradients (only data
    gradients).  Used for the discriminator pass of the generator update: the reference computes a
    D weight gradient there that ``net_d.zero_grad()`` discards before any use (train.py:58,107)."""
    global _skip_param_grads
    prev, _skip_param_grads = _skip_param_grads, True
    try:
        yield
    finally:
        _skip_param_grads = prev


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Named intermediate activations for per-layer parity checks (tests): inside ``record_taps`` the
# generator records its NHWC bf16 block outputs (``retain_grad`` so that backward leaves the
# activation gradient on them).  Outside the context ``tap`` is a no-op.
_taps = [None]


@contextlib.contextmanager
def record_taps():
    prev, _taps[0] = _taps[0], {}
    try:
        yield _taps[0]
    finally:
        _taps[0] = prev


def tap(name: str, x: torch.Tensor) -> torch.Tensor:
    if _taps[0] is not None:
        if x.requires_grad:
            x.retain_grad()
        _taps[0][name] = x
    return x


_stats_rows = [0]


def stats_rows() -> int:
    """Rows of a conv ``stats`` buffer: one partial-sum row per persistent CTA (SM count)."""
    if not _stats_rows[0]:
        _stats_rows[0] = int(query("sisr_stats_rows"))
    return _stats_rows[0]


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.SisrError(f"{what}: tensor is on {t.device}; the sisr_b200 operators run only on "
                             "CUDA (sm_100a) - there is no CPU path")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's stream; nn.DataParallel replicas (config.py:114-118) run in
        # threads on other devices and would also share this module's per-step state - use one process per
        # GPU (parallel.init_distributed / GradSync) instead
        raise _lib.SisrError(f"{what}: tensor is on {t.device} but the current device is cuda:"
                             f"{torch.cuda.current_device()}; single-process multi-device execution "
                             "(nn.DataParallel) is not supported - run one process per GPU")


_dist_group = None


def set_sync_group(group):
    """Process group used for SyncBN statistic reductions (None = single process)."""
    global _dist_group
    _dist_group = group


_peer = None


def set_peer_exchange(px):
    """parallel.PeerExchange used for SyncBN (None: all-reduce through the process group)."""
    global _peer
    _peer = px


class _ZeroArena:
    """One zero-filled fp32 buffer per step for the many tiny accumulators (BN backward sums, bias /
    slope gradients): a single memset at the start of the step instead of ~75 fill kernels.  Every
    slice is handed out at most once between two ``reset`` calls; without a trainer calling
    ``begin_step`` the arena is never armed and callers get ``torch.zeros``."""

    FLOATS = 1 << 19

    def __init__(self):
        self.buf, self.off = None, 0

    def reset(self, dev):
        if self.buf is None or self.buf.device != dev:
            self.buf = torch.zeros(self.FLOATS, dtype=torch.float32, device=dev)
        else:
            self.buf.zero_()
        self.off = 0

    def take(self, n, dev):
        if self.buf is None or self.buf.device != dev or self.off + n > self.FLOATS:
            return torch.zeros(n, dtype=torch.float32, device=dev)
        out = self.buf[self.off:self.off + n]
        self.off += _round_up(n, 4)
        return out


_arena = _ZeroArena()

# Weight gradients on a side stream (enabled by the trainer): the wgrad + finish kernels of a conv only
# feed the optimizer, so they run concurrently with the rest of the backward pass (dgrad / BatchNorm
# chain) instead of on its critical path.  In this mode the conv operators return no weight / bias
# gradient to autograd (which would read or clone the tensors on the main stream while the side
# stream still writes them); the gradients are attached to ``param.grad`` when the streams are joined
# at the end of the backward pass, and the parameters' post-accumulate hooks (gradient sync) are
# fired then.  A parameter used twice in one backward pass (the two discriminator passes of the D
# update) accumulates in the kernel (second launch: accumulate = 1).
_async = {"on": False, "stream": None, "pending": {}, "keep": [], "uses": {}, "track": False, "ready": []}
_ASYNC_LAG = 4      # a finished parameter is handed over once this many later launches are queued behind it


def _wgrad_stream():
    if _async["stream"] is None:
        _async["stream"] = torch.cuda.Stream()
    return _async["stream"]


_grad_ready_cb = [None]


def set_grad_ready_callback(fn):
    """``fn(param)`` is called whenever a side-stream weight gradient has been attached to ``param.grad``
    (those gradients do not pass through autograd's accumulation, so its post-accumulate hooks never
    see them).  parallel.GradSync registers itself here; None removes the callback."""
    _grad_ready_cb[0] = fn


_before_join_cb = [None]


def set_before_join_callback(fn):
    """``fn()`` runs at the end of a backward pass, after the current stream has waited for the weight-gradient
    stream and before the gradients still pending there are handed over (parallel.GradSync: see its ``drain``)."""
    _before_join_cb[0] = fn


def _deliver(entry):
    weight, bias, dw, db, want_w, want_b = entry
    for p, g, want in ((weight, dw, want_w), (bias, db, want_b)):
        if not want or p is None:
            continue
        p.grad = g if p.grad is None else p.grad + g
        if _grad_ready_cb[0] is not None:
            _grad_ready_cb[0](p)


def _deliver_ready(keep_last: int):
    """Hand over the parameters whose last launch is at least ``keep_last`` launches old: the main
    stream waits for the event recorded behind that launch (normally long finished), the gradient
    lands in ``param.grad`` and the gradient-sync hooks fire while backward is still running."""
    ready = _async["ready"]
    while len(ready) > keep_last:
        key, ev = ready.pop(0)
        entry = _async["pending"].pop(key, None)
        if entry is not None:
            torch.cuda.current_stream().wait_event(ev)
            _deliver(entry)


def join_wgrad(end_of_backward: bool = True):
    """Make the current stream wait for the outstanding side-stream weight gradients;
    ``end_of_backward``: also hand the finished gradients to their parameters."""
    if _async["stream"] is not None and (_async["keep"] or _async["pending"]):
        torch.cuda.current_stream().wait_stream(_async["stream"])
    _async["keep"].clear()
    if not end_of_backward:
        return

    _async["ready"].clear()
    pending, _async["pending"] = _async["pending"], {}
    # (only with gradients to hand over: begin_step() calls this too, possibly as the first thing of a CUDA-graph
    # capture, where a wait on the uncaptured all-reduces of the preceding eager steps would be illegal)
    if pending and _before_join_cb[0] is not None:
        _before_join_cb[0]()
    for entry in pending.values():
        _deliver(entry)


@contextlib.contextmanager
def async_weight_grads(on: bool = True):
    prev, _async["on"] = _async["on"], on
    try:
        yield
    finally:
        join_wgrad()
        _async["on"] = prev


_sync_scope = [False]


@contextlib.contextmanager
def sync_bn_scope():
    """SyncBN statistics are exchanged only inside this scope, which the trainer opens around a step that
    EVERY rank executes.  A forward pass outside it (one rank rendering a preview in train mode, as the
    reference's save_curr_vis does, utils.py:53-56) normalises with local statistics instead of
    desynchronising the per-slot epochs of the peer exchange - which used to end in a 10 s spin and a trap."""
    prev, _sync_scope[0] = _sync_scope[0], True
    try:
        yield
    finally:
        _sync_scope[0] = prev


def begin_step(device=None, track_weight_uses: bool = False):
    """Called by the trainer at the start of every step (all ranks): resets per-step numbering and
    re-zeroes the accumulator arena."""
    _colsum_cache.clear()
    _relu_outputs.clear()
    _premasked.clear()
    join_wgrad()
    _async["uses"].clear()
    _async["track"] = track_weight_uses
    if _peer is not None:
        _peer.reset()
    if device is not None and torch.device(device).type == "cuda":
        _arena.reset(torch.device(device))


def _world():
    import torch.distributed as dist
    if _dist_group is None or not _sync_scope[0] or not dist.is_initialized():
        return 1
    return dist.get_world_size(_dist_group)


def _all_reduce(t):
    import torch.distributed as dist
    dist.all_reduce(t, group=_dist_group)


# ----------------------------------------------------------------------------- layout
class ToNHWC(torch.autograd.Function):
    """NCHW fp32 (module boundary) -> NHWC bf16."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "ToNHWC")
        x = x.contiguous().float()
        n, c, h, w = x.shape
        y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
        call("sisr_nchw_f32_to_nhwc_bf16", x, y, n, c, h, w, _stream())
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        n, h, w, c = gy.shape
        gx = torch.empty((n, c, h, w), dtype=torch.float32, device=gy.device)
        call("sisr_nhwc_bf16_to_nchw_f32", gy, gx, n, c, h, w, _stream())
        return gx


class ToNCHW(torch.autograd.Function):
    """NHWC bf16 -> NCHW fp32 (module boundary)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        n, h, w, c = x.shape
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        call("sisr_nhwc_bf16_to_nchw_f32", x, y, n, c, h, w, _stream())
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous().float()
        n, c, h, w = gy.shape
        gx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=gy.device)
        call("sisr_nchw_f32_to_nhwc_bf16", gy, gx, n, c, h, w, _stream())
        return gx


# ----------------------------------------------------------------------------- batched weight preparation
class _SnLayerC(ctypes.Structure):
    """struct sisr_sn_layer"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("w", "u", "v", "t", "s", "sigma", "u_saved", "v_saved")] + \
               [(n, ctypes.c_int) for n in ("cout", "k", "training", "reserved")]


class _PrepLayerC(ctypes.Structure):
    """struct sisr_prep_layer"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("w", "sigma", "bias", "wf", "wd", "bias_perm")] + \
               [(n, ctypes.c_int) for n in ("cout", "cin", "k", "ps_r")]


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def prepare_convs(convs, first_needs_dx: bool):
    """Spectral-norm power iteration + bf16 weight re-layout for every conv of a network in one
    launch chain (3 + 1 kernels instead of 4 per layer).  ``convs``: the network's ``SNConv2d``
    modules in forward order.  Each module receives ``_prep`` = (wf, wd, sigma, u_saved, v_saved,
    bias_used), consumed by its next ``run``.  All outputs of one call live in two fresh buffers, so
    that several forward calls before a backward (the three discriminator passes of a step) keep
    their own sigma / weights, as the reference's per-call power iteration requires."""
    if not convs:
        return
    dev = convs[0].master_weight.device
    _require_cuda(convs[0].master_weight, "prepare_convs")
    st = _stream()
    f_off, h_off, plan = 0, 0, []
    for i, c in enumerate(convs):
        w = c.master_weight
        cout, cin, k, _ = w.shape
        K = cin * k * k
        need_dx = first_needs_dx or i > 0
        e = {"conv": c, "cout": cout, "cin": cin, "k": k, "K": K, "ps_r": getattr(c, "ps_r", 0)}
        if c.sn:
            for name, n in (("sigma", 4), ("t", K), ("s", cout), ("u_saved", cout), ("v_saved", K)):
                e[name] = (f_off, n)
                f_off += _round_up(n, 4)
        if e["ps_r"] == 2:
            e["bias_perm"] = (f_off, cout)
            f_off += _round_up(cout, 4)
        e["wf"] = (h_off, cout * K)
        h_off += _round_up(cout * K, 64)
        if need_dx:
            e["wd"] = (h_off, cout * K)
            h_off += _round_up(cout * K, 64)
        plan.append(e)
    fbuf = torch.empty(max(f_off, 4), dtype=torch.float32, device=dev)
    hbuf = torch.empty(max(h_off, 64), dtype=torch.bfloat16, device=dev)

    def fv(e, name):
        if name not in e:
            return None
        o, n = e[name]
        return fbuf[o:o + n]

    sn_rows, prep_rows = [], []
    for e in plan:
        c = e["conv"]
        w = c.master_weight
        cout, cin, k = e["cout"], e["cin"], e["k"]
        sigma = None
        if c.sn:
            sigma = fv(e, "sigma")[:1]
            sn_rows.append(_SnLayerC(w.data_ptr(), c.weight_u.data_ptr(), c.weight_v.data_ptr(),
                                     fv(e, "t").data_ptr(), fv(e, "s").data_ptr(), sigma.data_ptr(),
                                     fv(e, "u_saved").data_ptr(), fv(e, "v_saved").data_ptr(),
                                     cout, e["K"], 1 if c.training else 0, 0))
        o, n = e["wf"]
        wf = hbuf[o:o + n].view(cout, k, k, cin)
        wd = None
        if "wd" in e:
            o, n = e["wd"]
            wd = hbuf[o:o + n].view(cin, k, k, cout)
        bias_perm = fv(e, "bias_perm")
        prep_rows.append(_PrepLayerC(w.data_ptr(), sigma.data_ptr() if sigma is not None else None,
                                     c.bias.data_ptr(), wf.data_ptr(), wd.data_ptr() if wd is not None else None,
                                     bias_perm.data_ptr() if bias_perm is not None else None,
                                     cout, cin, k, e["ps_r"]))
        c._prep = (wf, wd, sigma, fv(e, "u_saved"), fv(e, "v_saved"),
                   bias_perm if bias_perm is not None else c.bias)
    if sn_rows:
        arr = (_SnLayerC * len(sn_rows))(*sn_rows)
        _lib.LAUNCHES[0] += 3 * ((len(sn_rows) + 39) // 40) - 1
        call("sisr_sn_power_iteration_batched", ctypes.addressof(arr), len(sn_rows), SN_EPS, st)
    arr = (_PrepLayerC * len(prep_rows))(*prep_rows)
    _lib.LAUNCHES[0] += (len(prep_rows) + 39) // 40 - 1
    call("sisr_weight_prep_batched", ctypes.addressof(arr), len(prep_rows), st)


# ----------------------------------------------------------------------------- convolution
@dataclass(frozen=True)
class ConvCfg:
    stride: int = 1
    pad: int = 1
    act: int = ACT_NONE
    leaky_slope: float = LEAKY_SLOPE
    ps_r: int = 0              # 2: PixelShuffle(2) folded into the store
    want_stats: bool = False   # BN batch statistics from the conv epilogue
    training: bool = True      # spectral-norm power iteration on/off
    out_nchw_f32: bool = False  # edge layer: write fp32 NCHW (+tanh) for the module boundary
    sync_wgrad: bool = False   # weight gradient returned through autograd even in side-stream mode (the weight
                               # is a derived tensor, e.g. zero-padded, whose gradient must flow on)


def _desc(x_shape, cout, k, cfg: ConvCfg) -> ConvDesc:
    n, h, w, cin = x_shape
    oh = (h + 2 * cfg.pad - k) // cfg.stride + 1
    ow = (w + 2 * cfg.pad - k) // cfg.stride + 1
    return ConvDesc(n, h, w, cin, oh, ow, cout, k, cfg.stride, cfg.pad, cfg.ps_r)


class Conv2dFn(torch.autograd.Function):
    """[spectral norm ->] conv2d + bias [+ PReLU/LeakyReLU/ReLU/Tanh] [+ PixelShuffle(2) store]
    [+ BN statistics].  ``weight`` is the fp32 master weight ([Cout,Cin,k,k]; ``weight_orig`` when
    ``u``/``v`` are given).  Returns (y, stats)."""

    @staticmethod
    def forward(ctx, x, weight, bias, u, v, slope, cfg: ConvCfg, prepared=None):
        _require_cuda(x, "Conv2dFn")
        x = x.contiguous()
        dev = x.device
        cout, cin, k, _ = weight.shape
        d = _desc(x.shape, cout, k, cfg)
        st = _stream()
        sigma = saved_u = saved_v = None
        if u is not None and (prepared is None or len(prepared) == 2):
            sigma = torch.empty(1, dtype=torch.float32, device=dev)
            ws = torch.empty(query("sisr_sn_workspace_floats", cout, cin * k * k), dtype=torch.float32,
                             device=dev)
            call("sisr_sn_power_iteration", weight, u, v, sigma, cout, cin * k * k,
                 1 if cfg.training else 0, SN_EPS, ws, st)
            saved_u, saved_v = u.clone(), v.clone()
        need_dx = x.requires_grad
        bias_used = bias
        if prepared is not None and len(prepared) == 2:   # frozen weights prepared once (MaskedVGG)
            wf, wd = prepared
        elif prepared is not None:                         # ops.prepare_convs (whole network at once)
            wf, wd, sigma, saved_u, saved_v, bias_used = prepared
        else:
            wf = torch.empty((cout, k, k, cin), dtype=torch.bfloat16, device=dev)
            wd = torch.empty((cin, k, k, cout), dtype=torch.bfloat16, device=dev) if need_dx else None
            if cfg.ps_r == 2:
                bias_used = torch.empty_like(bias)
            call("sisr_weight_prep", weight, sigma, bias, wf, wd, bias_used if cfg.ps_r == 2 else None,
                 cout, cin, k, cfg.ps_r, st)
        stats = torch.empty((stats_rows(), 2 * cout), dtype=torch.float32, device=dev) if cfg.want_stats else None
        if cfg.out_nchw_f32:
            y = torch.empty((d.n, cout, d.oh, d.ow), dtype=torch.float32, device=dev)
            call("sisr_conv_fprop", d, x, wf, bias_used, cfg.act, cfg.leaky_slope, slope, None, y, None, st)
        else:
            if cfg.ps_r == 2:
                y = torch.empty((d.n, d.oh * 2, d.ow * 2, cout // 4), dtype=torch.bfloat16, device=dev)
            else:
                y = torch.empty((d.n, d.oh, d.ow, cout), dtype=torch.bfloat16, device=dev)
            call("sisr_conv_fprop", d, x, wf, bias_used, cfg.act, cfg.leaky_slope, slope, y, None, stats, st)
        ctx.cfg, ctx.d = cfg, d
        ctx.weight_ref, ctx.bias_ref = weight, bias
        src = _relu_outputs.pop(x.data_ptr(), None)      # consumed by the one conv that reads this tensor
        ctx.input_is_relu = bool(src is not None and src.shape == x.shape and need_dx and not cfg.out_nchw_f32
                                 and query("sisr_conv_dgrad_fuses_mask", d))
        if (cfg.act == ACT_RELU and need_dx and not cfg.out_nchw_f32 and cfg.ps_r != 2 and
                (_skip_param_grads or not (weight.requires_grad or bias.requires_grad))):
            # entries are removed when consumed; the ones never consumed (a ReLU output that feeds a max-pool)
            # are evicted oldest-first, so that callers that never reach begin_step() (the drop-in mode of
            # INTEGRATION.md) pin at most 16 activations instead of growing by four per VGG forward
            while len(_relu_outputs) >= 16:
                del _relu_outputs[next(iter(_relu_outputs))]
            _relu_outputs[y.data_ptr()] = y
        if _async["track"] and not _skip_param_grads and (weight.requires_grad or bias.requires_grad):
            k_ = weight.data_ptr()
            _async["uses"][k_] = _async["uses"].get(k_, 0) + 1
        ctx.has_sn = u is not None
        ctx.has_slope = slope is not None
        ctx.skip_params = _skip_param_grads
        ctx.save_for_backward(x, weight, wf, wd, y if cfg.act != ACT_NONE else None, sigma, saved_u,
                              saved_v, slope)
        if stats is not None:
            ctx.mark_non_differentiable(stats)
        ctx.set_materialize_grads(False)      # no zero tensor for the gradient of ``stats`` (one fill kernel per layer)
        return y, stats

    @staticmethod
    def backward(ctx, gy, _gstats):
        x, weight, wf, wd, y, sigma, u, v, slope = ctx.saved_tensors
        cfg, d = ctx.cfg, ctx.d
        dev = x.device
        st = _stream()
        cout, cin, k, _ = weight.shape
        if gy is None:                        # only the statistics were used downstream
            return (None,) * 8
        gy = gy.contiguous()
        dslope = colsum = None
        if cfg.out_nchw_f32:
            dpre = torch.empty((d.n, d.oh, d.ow, cout), dtype=torch.bfloat16, device=dev)
            if cfg.act == ACT_TANH:
                call("sisr_tanh_bwd_nchw_to_nhwc", gy.float(), y, dpre, d.n, cout, d.oh, d.ow, st)
            else:
                call("sisr_nchw_f32_to_nhwc_bf16", gy.float(), dpre, d.n, cout, d.oh, d.ow, st)
        elif cfg.act == ACT_RELU and _premasked.pop(gy.data_ptr(), None) is not None:
            dpre = gy                      # the consumer's dgrad already applied this layer's ReLU mask
        elif cfg.act != ACT_NONE:
            dpre = torch.empty_like(gy)
            want_ds = cfg.act == ACT_PRELU and ctx.needs_input_grad[5] and not ctx.skip_params
            want_db = ctx.needs_input_grad[2] and not ctx.skip_params and cfg.ps_r != 2
            c_last = gy.shape[-1]
            if want_ds or want_db:
                zbuf = _arena.take(1 + c_last, dev)
                dslope = zbuf[:1] if want_ds else None
                colsum = zbuf[1:] if want_db else None
            call("sisr_act_bwd", gy, y, cfg.act, cfg.leaky_slope, slope, dpre, dslope, colsum,
                 gy.numel() // c_last, c_last, st)
        else:
            dpre = gy
            hit = _colsum_cache.pop(gy.data_ptr(), None)
            if hit is not None and hit[0].numel() == gy.numel() and cfg.ps_r != 2:
                colsum = hit[1]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if ctx.input_is_relu:
                call("sisr_conv_dgrad_masked", d, dpre, wf, wd, dx, x, 0.0, st)
                while len(_premasked) >= 16:  # callers that never reach begin_step(): oldest first
                    del _premasked[next(iter(_premasked))]
                _premasked[dx.data_ptr()] = dx
            else:
                call("sisr_conv_dgrad", d, dpre, wf, wd, dx, st)
        dw = db = None
        if (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and not ctx.skip_params:
            nbytes = query("sisr_conv_wgrad_fused_workspace_bytes", d)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            sig = sigma if ctx.has_sn else None
            if _async["on"] and not cfg.sync_wgrad:
                key = weight.data_ptr()
                prev = _async["pending"].get(key)
                if prev is None:
                    dw = torch.empty_like(weight)
                    db = torch.empty(cout, dtype=torch.float32, device=dev)
                    _async["pending"][key] = (ctx.weight_ref, ctx.bias_ref, dw, db, ctx.needs_input_grad[1],
                                              ctx.needs_input_grad[2])
                else:
                    dw, db = prev[2], prev[3]
                side = _wgrad_stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    call("sisr_conv_wgrad_fused", d, x, dpre, weight, u, v, sig, dw, colsum, db,
                         0 if prev is None else 1, ws, side.cuda_stream)
                # alive until the streams are joined: everything the side-stream kernels read.  u / v / sigma are
                # slices of prepare_convs' per-forward buffer, which is otherwise freed as soon as the last conv of
                # the network has been back-propagated - while the finish kernels of the last layers are still
                # queued - and handed to the next allocation of whichever stream it came from
                _async["keep"].append((ws, x, dpre, colsum, u, v, sig))
                dw = db = None            # delivered to param.grad by _deliver_ready() / join_wgrad()
                left = _async["uses"].get(key)
                if left is not None:
                    _async["uses"][key] = left - 1
                    if left - 1 == 0:     # every forward use of this parameter has been back-propagated
                        ev = torch.cuda.Event()
                        ev.record(side)
                        _async["ready"].append((key, ev))
                        _deliver_ready(_ASYNC_LAG)
            else:
                dw = torch.empty_like(weight)
                db = torch.empty(cout, dtype=torch.float32, device=dev)
                call("sisr_conv_wgrad_fused", d, x, dpre, weight, u, v, sig, dw, colsum, db, 0, ws, st)
            if not ctx.needs_input_grad[1]:
                dw = None
            if not ctx.needs_input_grad[2]:
                db = None
        return dx, dw, db, None, None, dslope, None, None


# ----------------------------------------------------------------------------- BatchNorm (+act, +residual)
@dataclass(frozen=True)
class BnCfg:
    act: int = ACT_NONE
    leaky_slope: float = LEAKY_SLOPE
    training: bool = True
    momentum: float = BN_MOMENTUM
    eps: float = BN_EPS
    sync: bool = True          # reduce statistics over the data-parallel group when one is set


class BnActFn(torch.autograd.Function):
    """out = act(BatchNorm(y)) [+ residual] on NHWC bf16; ``stats`` = per-channel {sum, sum sq} of
    ``y`` from the producing conv's epilogue (computed here when None)."""

    @staticmethod
    def forward(ctx, y, stats, gamma, beta, running_mean, running_var, nbt, residual, slope, cfg: BnCfg):
        y = y.contiguous()
        dev = y.device
        c = y.shape[-1]
        rows = y.numel() // c
        st = _stream()
        count = float(rows)
        if cfg.training:
            if stats is None:
                stats = torch.empty((1, 2 * c), dtype=torch.float32, device=dev)
                call("sisr_bn_stats", y, rows, c, stats, st)
            world = _world() if cfg.sync else 1
            if world > 1:
                count *= world
                if _peer is None:
                    stats = stats.sum(dim=0, keepdim=True) if stats.shape[0] > 1 else stats.clone()
                    _all_reduce(stats)
        else:
            stats = torch.zeros((1, 2 * c), dtype=torch.float32, device=dev)
        aux = torch.empty((4, c), dtype=torch.float32, device=dev)  # scale, shift, mean, invstd
        if cfg.training and cfg.sync and _world() > 1 and _peer is not None:
            # global-batch statistics: partial rows + NVLink peer exchange + finalize in one kernel
            call("sisr_bn_finalize_sync", _peer.bases, _peer.rank, _peer.world, _peer.next_slot(), stats,
                 stats.shape[0], count, gamma, beta, running_mean, running_var, nbt, cfg.momentum,
                 cfg.eps, aux[0], aux[1], aux[2], aux[3], c, st)
        else:
            call("sisr_bn_finalize", stats, stats.shape[0], count, gamma, beta, running_mean, running_var, nbt,
                 cfg.momentum, cfg.eps, 1 if cfg.training else 0, aux[0], aux[1], aux[2], aux[3], c, st)
        out = torch.empty_like(y)
        if residual is not None:
            residual = residual.contiguous()
        call("sisr_bn_apply", y, aux[0], aux[1], cfg.act, cfg.leaky_slope, slope, residual, out, rows, c, st)
        ctx.cfg, ctx.count = cfg, count
        ctx.has_residual = residual is not None
        ctx.skip_params = _skip_param_grads
        ctx.save_for_backward(y, aux, slope)
        return out

    @staticmethod
    def backward(ctx, gout):
        y, aux, slope = ctx.saved_tensors
        cfg = ctx.cfg
        dev = y.device
        c = y.shape[-1]
        rows = y.numel() // c
        st = _stream()
        gout = gout.contiguous()
        zbuf = _arena.take(3 * c + 2, dev)
        sums = zbuf[:2 * c + 1]
        colsum = zbuf[2 * c + 1:3 * c + 1] if not ctx.skip_params else None
        local = sums
        synced = cfg.training and cfg.sync and _world() > 1
        if synced and _peer is not None:
            # local sums + NVLink exchange in one launch (the last CTA to finish runs the exchange)
            red = torch.empty(2 * c + 1, dtype=torch.float32, device=dev)
            call("sisr_bn_bwd_reduce_sync", _peer.bases, _peer.rank, _peer.world, _peer.next_slot(), gout, y,
                 aux[2], aux[3], aux[0], aux[1], cfg.act, cfg.leaky_slope, slope, sums, red, zbuf[3 * c + 1:],
                 rows, c, st)
        else:
            call("sisr_bn_bwd_reduce", gout, y, aux[2], aux[3], aux[0], aux[1], cfg.act, cfg.leaky_slope,
                 slope, sums, rows, c, st)
            if synced:
                sums = sums.clone()
                _all_reduce(sums)
            red = sums if cfg.training else torch.zeros_like(sums)   # eval: statistics are constants
        dy = None
        if ctx.needs_input_grad[0]:
            dy = torch.empty_like(y)
            call("sisr_bn_bwd_apply", gout, y, aux[2], aux[3], aux[0], aux[1], cfg.act, cfg.leaky_slope,
                 slope, red, ctx.count, dy, colsum, rows, c, st)
            if colsum is not None:
                # per-channel sum of dy = bias gradient of the producing conv; handed to its backward
                while len(_colsum_cache) >= 16:
                    del _colsum_cache[next(iter(_colsum_cache))]
                _colsum_cache[dy.data_ptr()] = (dy, colsum)
        dgamma = dbeta = dslope = None
        if not ctx.skip_params:
            if ctx.needs_input_grad[2]:
                dgamma = local[c:2 * c].clone()
            if ctx.needs_input_grad[3]:
                dbeta = local[:c].clone()
            if slope is not None and ctx.needs_input_grad[8]:
                dslope = local[2 * c:2 * c + 1].clone()
        dres = gout if (ctx.has_residual and ctx.needs_input_grad[7]) else None
        return dy, None, dgamma, dbeta, None, None, None, dres, dslope, None


# ----------------------------------------------------------------------------- pooling
class MaxPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        n, h, w, c = x.shape
        y = torch.empty((n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
        call("sisr_maxpool2_fwd", x, y, n, h, w, c, _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        n, h, w, c = x.shape
        dx = torch.empty_like(x)
        call("sisr_maxpool2_bwd", x, gy.contiguous(), dx, n, h, w, c, _stream())
        return dx


# ----------------------------------------------------------------------------- discriminator head
class DHeadFn(torch.autograd.Function):
    """flatten (reference (c,h,w) order) -> Linear -> LeakyReLU -> Linear -> Sigmoid; x is NHWC bf16."""

    @staticmethod
    def forward(ctx, x, w0, b0, w2, b2):
        x = x.contiguous()
        dev = x.device
        n, h, w, c = x.shape
        fc_in, fc_mid = h * w * c, w0.shape[0]
        st = _stream()
        xf = torch.empty((n, fc_in), dtype=torch.bfloat16, device=dev)
        call("sisr_transpose_bf16", x, xf, n, h * w, c, st)
        hbuf = torch.empty((n, fc_mid), dtype=torch.float32, device=dev)
        p = torch.empty((n, 1), dtype=torch.float32, device=dev)
        call("sisr_dhead_forward", xf, w0, b0, w2, b2, LEAKY_SLOPE, hbuf, p, n, fc_in, fc_mid, st)
        ctx.shape = (n, h, w, c)
        ctx.skip_params = _skip_param_grads
        ctx.save_for_backward(xf, w0, w2, hbuf, p)
        return p

    @staticmethod
    def backward(ctx, gp):
        xf, w0, w2, hbuf, p = ctx.saved_tensors
        n, h, w, c = ctx.shape
        dev = xf.device
        fc_in, fc_mid = xf.shape[1], w0.shape[0]
        st = _stream()
        need_w = any(ctx.needs_input_grad[1:]) and not ctx.skip_params
        dh = torch.empty((n, fc_mid), dtype=torch.float32, device=dev)
        dw0 = torch.empty_like(w0) if need_w else None
        db0 = torch.empty(fc_mid, dtype=torch.float32, device=dev)
        dw2 = torch.empty((1, fc_mid), dtype=torch.float32, device=dev)
        db2 = torch.empty(1, dtype=torch.float32, device=dev)
        dxf = torch.empty((n, fc_in), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        call("sisr_dhead_backward", xf, w0, w2, hbuf, p, gp.contiguous().float(), LEAKY_SLOPE, dh, dw0,
             db0, dw2, db2, dxf, n, fc_in, fc_mid, 1 if need_w else 0, st)
        dx = None
        if dxf is not None:
            dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dev)
            call("sisr_nchw_f32_to_nhwc_bf16", dxf, dx, n, c, h, w, st)
        if not need_w:
            return dx, None, None, None, None
        return dx, dw0, db0, dw2, db2


# ----------------------------------------------------------------------------- losses
class BCEFn(torch.autograd.Function):
    """nn.BCELoss(reduction='mean') of probabilities ``p`` against a constant target."""

    @staticmethod
    def forward(ctx, p, target: float):
        p = p.contiguous().float().view(-1)
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        mean_p = torch.empty(1, dtype=torch.float32, device=p.device)
        call("sisr_bce_fwd", p, p.numel(), float(target), loss, mean_p, _stream())
        ctx.target = float(target)
        ctx.save_for_backward(p)
        ctx.mark_non_differentiable(mean_p)
        return loss.view(()), mean_p.view(())

    @staticmethod
    def backward(ctx, gl, _gm):
        (p,) = ctx.saved_tensors
        dp = torch.empty_like(p)
        call("sisr_bce_bwd", p, p.numel(), ctx.target, gl.contiguous().float().view(1), dp, _stream())
        return dp, None


class MSEFn(torch.autograd.Function):
    """torch.mean((a - b)**2) on fp32 tensors of equal shape (train.py:186)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous().float(), b.contiguous().float()
        n = a.numel()
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        call("sisr_mse_fwd", a, b, n, 1.0 / n, loss, _stream())
        ctx.save_for_backward(a, b)
        return loss.view(())

    @staticmethod
    def backward(ctx, gl):
        a, b = ctx.saved_tensors
        n = a.numel()
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        call("sisr_mse_bwd", a, b, n, 1.0 / n, gl.contiguous().float().view(1), ga, gb, _stream())
        return ga, gb


class PixelShuffle2Fn(torch.autograd.Function):
    """nn.PixelShuffle(2) on NHWC bf16: [N,H,W,4C] -> [N,2H,2W,C] (stand-alone kernel)."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "PixelShuffle2Fn")
        x = x.contiguous()
        n, h, w, c4 = x.shape
        if c4 % 4:
            raise _lib.SisrError("PixelShuffle(2) needs a channel count divisible by 4")
        y = torch.empty((n, 2 * h, 2 * w, c4 // 4), dtype=x.dtype, device=x.device)
        call("sisr_pixel_shuffle2", x, y, n, h, w, c4 // 4, 0, _stream())
        ctx.shape = (n, h, w, c4)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, h, w, c4 = ctx.shape
        gx = torch.empty((n, h, w, c4), dtype=gy.dtype, device=gy.device)
        call("sisr_pixel_shuffle2", gy.contiguous(), gx, n, h, w, c4 // 4, 1, _stream())
        return gx


class LrFromHrFn(torch.autograd.Function):
    """utils.lr_from_hr: bicubic (align_corners=True) down-sampling + clamp to [-1, 1], NCHW fp32."""

    @staticmethod
    def forward(ctx, hr, size):
        _require_cuda(hr, "lr_from_hr")
        hr = hr.contiguous().float()
        n, c, h, w = hr.shape
        oh, ow = int(size[0]), int(size[1])
        lr = torch.empty((n, c, oh, ow), dtype=torch.float32, device=hr.device)
        call("sisr_lr_from_hr", hr, lr, n, c, h, w, oh, ow, _stream())
        ctx.save_for_backward(hr)
        ctx.size = (oh, ow)
        return lr

    @staticmethod
    def backward(ctx, glr):
        (hr,) = ctx.saved_tensors
        n, c, h, w = hr.shape
        oh, ow = ctx.size
        dhr = torch.empty_like(hr)
        call("sisr_lr_from_hr_bwd", hr, glr.contiguous().float(), dhr, n, c, h, w, oh, ow, _stream())
        return dhr, None


def bce_loss(p: torch.Tensor, target: float):
    return BCEFn.apply(p, target)


def mse_loss(a: torch.Tensor, b: torch.Tensor):
    return MSEFn.apply(a, b)
