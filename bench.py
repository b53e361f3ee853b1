"""Benchmark of the SRGAN x4 training step (BASELINE.json metric: HR patches / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

N = 1 workload = BASELINE.json configs[1]: full x4 step (G = GeneratorSuffix(Generator(16 blocks)),
D at 3x96x96, MaskedVGG54 content loss + adversarial loss), 96x96 HR patches, batch 64 per GPU,
bf16 storage / fp32 accumulate, synthetic U[-1,1] patches, random-init weights.  N > 1 (launched by
torchrun, one rank per GPU): same per-GPU batch (weak scaling), NCCL gradient all-reduce + SyncBN.

One JSON line on stdout (rank 0).  ``value`` = patches/s with the batch resident in HBM (whole step
replayed as one CUDA graph at N = 1); ``e2e`` = the same step fed from pinned HOST buffers with the
H2D copy of HR+LR and a D2H read of the three losses inside the timed region; ``roofline`` = the
dominant kernel (tcgen05 implicit-GEMM conv) timed alone (CUDA graph of back-to-back launches, CUDA
events) against the measured bf16 peak; ``roofline_hbm`` = the fused BN / PReLU kernels against the
measured HBM bandwidth; ``cpu_baseline`` = the CPU oracle port of the reference step on the box's host cores.
``--impl reference`` times the reference's own CPU implementation of the step (the oracle port:
/root/reference does not exist on the GPU box) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PATCH = 42.55e9          # 3 F_G + 8 F_D + 3 F_V at 96x96
# BASELINE.json configs[1] / [3] / [4]: algorithmic GFLOP per HR patch of the full step
CONFIGS = {
    "x4": {"hr": 96, "lr": 24, "n_suffix": 1, "frozen": False, "flop": 42.55e9, "batch": 64,
           "workload": "SRGAN x4 full training step (G 16 blocks + suffix, D @3x96x96, MaskedVGG54 content loss "
                       "+ adversarial loss), 96x96 HR patches"},
    "frozen": {"hr": 96, "lr": 24, "n_suffix": 1, "frozen": True, "flop": 38.66e9, "batch": 64,
               "workload": "progressive x4 = x2 generator wrapped by GeneratorSuffix(freeze_prefix, freeze_upscale, "
                           "freeze_end) (config.py:96): frozen 16-block trunk, spectral norm in G and D, D @3x96x96, "
                           "MaskedVGG54, 96x96 HR patches"},
    "x8": {"hr": 256, "lr": 32, "n_suffix": 2, "frozen": False, "flop": 280.75e9, "batch": 16,
           "workload": "progressive x8 supervised SRGAN: GeneratorSuffix(GeneratorSuffix(Generator)), 32x32 LR -> "
                       "256x256 HR patches, 256-channel pre-upsample maps, D @3x256x256 (138.9 M parameters), "
                       "MaskedVGG54"},
}
D_FEATS = [64, 64, 128, 128, 256, 256, 512, 512]
D_STRIDES = [1, 2, 1, 2, 1, 2, 1, 2]
VGG54 = 0b10000


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "hbm": p["hbm_gbs"], "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), [s.strip() for s in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in the wall-clock window [t0, t1] (the timed regions)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        return self.window(t0, t1)

    def window(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [s for (t, s) in self.samples if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.2)]
        sm = sorted(int(s[0]) for s in rows if s and s[0].isdigit())
        mx = [int(s[1]) for (_, s) in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in rows if len(s) >= 6
                          for i in range(4) if s[2 + i].lower().startswith("active")})
        pw = sorted(float(s[6]) for s in rows if len(s) >= 7 and s[6].replace(".", "", 1).isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None}


# ------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step_time(batch, steps, warmup, threads):
    """Seconds per step of the oracle port of train.train_loop's body on the host cores."""
    import torch
    from oracle import srgan_oracle as O
    from oracle import state_factory as S
    torch.set_num_threads(threads)
    g = S.generator_state(1, n_blocks=16, n_suffix=1)
    d = S.discriminator_state(2, (3, 96, 96), D_FEATS, D_STRIDES)
    v = S.vgg_state(3, VGG54)
    og = O.AdamState(O.trainable_names(g), 1e-5)
    od = O.AdamState(O.trainable_names(d), 1e-5)
    times = []
    for i in range(warmup + steps):
        hr = S.synthetic_hr(10 + i, batch, 96)
        lr = O.lr_from_hr(hr, (24, 24))
        t0 = time.perf_counter()
        O.train_step(g, d, v, hr, lr, d_strides=D_STRIDES, vgg_mask=VGG54, opt_g=og, opt_d=od)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


class _TimedBatches:
    """dataloader stand-in for the reference's train_loop: records the wall time of every fetch, so the
    time of step i is stamp[i+1] - stamp[i] without touching the reference's code."""

    def __init__(self, batches):
        self.batches, self.stamps = batches, []

    def __iter__(self):
        for b in self.batches:
            self.stamps.append(time.perf_counter())
            yield b

    def __len__(self):
        return len(self.batches)


def reference_train_loop_time(batch, steps, warmup, threads, budget_s=150.0):
    """Seconds per step of the reference's OWN train.train_loop (unmodified, from oracle/_ref) on the host
    cores: config 2 nets (G 16 blocks + suffix, D @96, MaskedVGG54, spectral norm on), synthetic batches.
    Returns (sec_per_step, steps_timed) or None when oracle/_ref is absent."""
    import torch
    from oracle import ref_harness as R
    from oracle import state_factory as S
    if R.reference_dir() is None:
        return None
    torch.set_num_threads(threads)
    mg, md, mc = R.import_reference()
    net_g = mg.GeneratorSuffix(mg.Generator(16, 64, 256, [2], use_sn=True))
    net_d = md.Discriminator((3, 96, 96), D_FEATS, D_STRIDES)
    ext = mc.MaskedVGG(VGG54)
    torch.nn.Module.load_state_dict(net_g, S.generator_state(1, n_blocks=16, n_suffix=1), strict=True)
    torch.nn.Module.load_state_dict(net_d, S.discriminator_state(2, (3, 96, 96), D_FEATS, D_STRIDES), strict=True)
    torch.nn.Module.load_state_dict(ext, S.vgg_state(3, VGG54), strict=True)
    # one probing step bounds the run: the whole arm must end within a few minutes
    probe = _TimedBatches([S.synthetic_hr(9, batch, 96), S.synthetic_hr(9, batch, 96)])
    R.run_train_loop(net_g, net_d, ext, probe, lr=1e-5, lr_size=24)
    t_probe = time.perf_counter() - probe.stamps[0]
    warmup = max(0, warmup - 1)
    steps = max(2, min(steps, int(budget_s / max(t_probe, 1e-3)) - warmup))
    data = _TimedBatches([S.synthetic_hr(10 + i, batch, 96) for i in range(warmup + steps + 1)])
    R.run_train_loop(net_g, net_d, ext, data, lr=1e-5, lr_size=24)
    st = data.stamps
    return (st[warmup + steps] - st[warmup]) / steps, steps


def run_reference(args):
    """CPU arm: the reference's own implementation of the step on the box's host cores.  With oracle/_ref
    present (made by __graft_entry__.build() in the dev container; it travels with the snapshot) this is the
    real, unmodified train.train_loop at the metric's batch 64 (kind "reference"); otherwise the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.batch or 64
    got = reference_train_loop_time(batch, args.steps, args.warmup, threads)
    if got is not None:
        sec, steps = got
        kind, warmup = "reference", max(1, args.warmup)
        sample = (f"{steps} step(s) of batch {batch} after {warmup} warm-up step(s): the unmodified reference "
                  f"train.train_loop (oracle/_ref), fp32, {threads} host threads")
        workload = "SRGAN x4 full training step (G+D+MaskedVGG54), 96x96 HR, CPU fp32, batch %d" % batch
    else:
        batch = 16                                   # bounded sample of the B=64 workload
        steps, warmup = min(args.steps, 3), min(args.warmup, 1)
        sec = cpu_reference_step_time(batch, steps, warmup, threads)
        kind = "port"
        sample = (f"{steps} step(s) of batch {batch} after {warmup} warm-up, oracle port of train.train_loop "
                  "body (oracle/_ref absent)")
        workload = ("SRGAN x4 full training step (G+D+MaskedVGG54), 96x96 HR, CPU fp32, "
                    f"sample batch {batch} of the batch-64 workload")
    value = batch / sec
    line = {
        "impl": "reference", "metric": "SRGAN x4 train-step HR patches/sec", "value": value,
        "unit": "patches/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": batch, "spectral_norm": True},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def build_trainer(dev, batch, world, grad_sync=None, config="x4"):
    import torch
    import sisr_b200 as m
    cfg = CONFIGS[config]
    torch.manual_seed(0)
    net_g = m.Generator(16, 64, 256, [2], use_sn=True)
    for i in range(cfg["n_suffix"]):
        last = i == cfg["n_suffix"] - 1
        if cfg["frozen"] and last:
            net_g = m.GeneratorSuffix(net_g, freeze_prefix=True, freeze_upscale=True, freeze_end=True)
        else:
            net_g = m.GeneratorSuffix(net_g)
    net_g = net_g.to(dev)
    net_d = m.Discriminator((3, cfg["hr"], cfg["hr"]), D_FEATS, D_STRIDES).to(dev)
    ext = m.MaskedVGG(VGG54).to(dev)
    if world > 1:
        from sisr_b200 import parallel
        for net in (net_g, net_d, ext):
            parallel.broadcast_module(net)
    # with a GradSync the trainer broadcasts weights + buffers and attaches both optimizers itself
    cfg_step = m.StepConfig(lr=1e-5, use_replay=False)
    if os.environ.get("SISR_OVERLAP") == "0":       # A-B timing: D(real) and MaskedVGG(fake) on the main stream
        cfg_step.overlap_d_real = cfg_step.overlap_fake_features = False
    return m.SRGANTrainer(net_g, net_d, ext, cfg_step, grad_sync=grad_sync)


DP_GOLDEN = os.path.join(ROOT, "profiles", "r2_dp_check_n1.json")


def dp_check(dev, rank, world, tr_timed):
    """Data-parallel correctness evidence carried by the bench line (all ranks call this):
      * ``weights_identical``: after the timed steps every rank holds bit-identical G / D weights, BN running
        statistics and spectral-norm vectors (an int64 checksum of the raw bits, all-gathered);
      * ``losses``: ONE step of a fresh, identically seeded trainer on a fixed 64-patch global batch sharded
        64/N per rank (SyncBN + gradient all-reduce => the single-process step on the whole batch), the three
        losses averaged over ranks, compared with the committed N = 1 values (profiles/r2_dp_check_n1.json,
        written by an N = 1 run of this same function) within 5e-3."""
    import torch
    import torch.distributed as dist
    from sisr_b200 import parallel
    import torch.nn.functional as F

    def checksum(nets):
        acc = torch.zeros((), dtype=torch.int64, device=dev)
        for net in nets:
            for t in list(net.parameters()) + list(net.buffers()):
                if t.dtype == torch.float32:
                    acc += t.detach().contiguous().view(torch.int32).to(torch.int64).sum()
        return acc
    mine = checksum((tr_timed.net_g, tr_timed.net_d)).reshape(1)
    sums = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(sums, mine)
    else:
        sums = [mine]
    identical = all(int(x) == int(sums[0]) for x in sums)
    gs = parallel.GradSync() if world > 1 else None
    tr = build_trainer(dev, 64 // world, world, gs)
    gen = torch.Generator().manual_seed(4242)
    hr_all = torch.rand((64, 3, 96, 96), generator=gen) * 2 - 1
    lr_all = F.interpolate(hr_all, (24, 24), mode="bicubic", align_corners=True).clamp(-1, 1)
    per = 64 // world
    sl = slice(rank * per, (rank + 1) * per)
    out = tr.step(hr_all[sl].to(dev), lr_all[sl].to(dev))
    losses = torch.stack([out["err_d"].reshape(()), out["err_g_adv"].reshape(()), out["err_g_cont"].reshape(())])
    if world > 1:
        dist.all_reduce(losses)
        losses /= world
    losses = [float(x) for x in losses]
    res = {"weights_identical": identical, "checksums": [int(x) for x in sums], "losses": losses,
           "global_batch": 64, "per_rank": per}
    if os.path.exists(DP_GOLDEN):
        want = json.load(open(DP_GOLDEN))["losses"]
        res["n1_losses"] = want
        rd = [abs(a - b) / abs(b) for a, b in zip(losses, want)]
        res["rel_diff_vs_n1"] = rd
        # err_d and err_g_cont are evaluated on the step's initial weights: 5e-3.  err_g_adv is evaluated AFTER
        # the discriminator's first Adam update of the same step, which is sign-like (every one of the 23.6 M
        # weights moves by +-lr whatever |g|), so the fp32 summation order of near-zero gradients shows in it:
        # 2e-2 (measured 5.8e-3 at N = 2; tools/multi_gpu_check.py compares the synchronised gradients directly)
        res["losses_match_n1"] = rd[0] < 5e-3 and rd[2] < 5e-3 and rd[1] < 2e-2
    if world == 1 and rank == 0 and os.environ.get("SISR_WRITE_DP_GOLDEN"):
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({"losses": losses, "how": "bench.py dp_check at N=1: one step, fresh seeded trainer, fixed "
                   "64-patch batch (seed 4242), lr 1e-5"}, open(os.path.join(ROOT, "gpurun_out", "r2_dp_check_n1.json"), "w"))
    return res


def _graph_time_us(launch, reps=12, replays=5):
    """Device time per launch: `reps` launches captured in one CUDA graph (no host gaps; the tensor maps
    are encoded at capture), timed with CUDA events on the replaying stream."""
    import torch
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(3):
            launch(i, side.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for i in range(reps):
            launch(i, st)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * replays)


def time_conv_kernel(dev, n, h, cin, cout, with_stats):
    """One tcgen05 conv launch through the C ABI, timed inside a CUDA graph of back-to-back launches that
    rotate over 3 input / output sets (the trunk tensors are L2-resident in the real step as well)."""
    import torch
    from sisr_b200 import _lib
    sets = []
    for _ in range(3):
        x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
        wt = (torch.randn(cout, 3, 3, cin, device=dev) * 0.02).to(torch.bfloat16)
        y = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
        stats = torch.empty(_lib.query("sisr_stats_rows"), 2 * cout, device=dev) if with_stats else None
        sets.append((x, wt, y, stats))
    bias = torch.zeros(cout, device=dev)
    d = _lib.ConvDesc(n, h, h, cin, h, h, cout, 3, 1, 1, 0)

    def launch(i, st):
        x, wt, y, stats = sets[i % 3]
        _lib.call("sisr_conv_fprop", d, x, wt, bias, 0, 0.0, None, y, None, stats, st)
    us = _graph_time_us(launch)
    flops = 2.0 * n * h * h * cout * 9 * cin
    return {"kernel": "conv3x3 %d->%d @%dx%d batch %d%s" % (cin, cout, h, h, n, " +BN stats" if with_stats else ""),
            "flops": flops, "ms": us * 1e-3, "tflops": flops / (us * 1e-6) / 1e12}


def time_conv_dgrad(dev, n, h, cin, cout, stride=1, ps=0):
    """Data-gradient launch of a conv layer through the C ABI (same graph timing)."""
    import torch
    from sisr_b200 import _lib
    oh = (h + 2 - 3) // stride + 1
    d = _lib.ConvDesc(n, h, h, cin, oh, oh, cout, 3, stride, 1, ps)
    sets = []
    for _ in range(3):
        wf = (torch.randn(cout, 3, 3, cin, device=dev) * 0.02).to(torch.bfloat16)
        wd = (torch.randn(cin, 3, 3, cout, device=dev) * 0.02).to(torch.bfloat16)
        shp = (n, oh * 2, oh * 2, cout // 4) if ps == 2 else (n, oh, oh, cout)
        dy = torch.randn(*shp, device=dev).to(torch.bfloat16)
        dx = torch.empty(n, h, h, cin, device=dev, dtype=torch.bfloat16)
        sets.append((wf, wd, dy, dx))

    def launch(i, st):
        wf, wd, dy, dx = sets[i % 3]
        _lib.call("sisr_conv_dgrad", d, dy, wf, wd, dx, st)
    us = _graph_time_us(launch)
    flops = 2.0 * n * oh * oh * cout * 9 * cin
    return {"kernel": "dgrad of conv3x3 %d->%d s%d @%dx%d" % (cin, cout, stride, h, h), "flops": flops,
            "ms": us * 1e-3, "tflops": flops / (us * 1e-6) / 1e12}


def time_dominant_kernel(dev, batch, frozen=False):
    """igemm_t_kernel (18.6 % of the step's kernel time, profiles/r1_step_launches_b64_final.csv) over
    ALL of its 93 launches per step: every distinct (shape, direction) it serves is timed and weighted by
    its launch count; achieved = sum of algorithmic FLOPs / sum of launch durations."""
    import torch
    from sisr_b200 import _lib

    def fwd(h, cin, cout, stats, stride=1):
        sets = []
        oh = (h + 2 - 3) // stride + 1
        d = _lib.ConvDesc(batch, h, h, cin, oh, oh, cout, 3, stride, 1, 0)
        for _ in range(3):
            x = torch.randn(batch, h, h, cin, device=dev).to(torch.bfloat16)
            wt = (torch.randn(cout, 3, 3, cin, device=dev) * 0.02).to(torch.bfloat16)
            y = torch.empty(batch, oh, oh, cout, device=dev, dtype=torch.bfloat16)
            st_ = torch.empty(_lib.query("sisr_stats_rows"), 2 * cout, device=dev) if stats else None
            sets.append((x, wt, y, st_))
        bias = torch.zeros(cout, device=dev)

        def launch(i, st):
            x, wt, y, st_ = sets[i % 3]
            _lib.call("sisr_conv_fprop", d, x, wt, bias, 0, 0.0, None, y, None, st_, st)
        us = _graph_time_us(launch)
        flops = 2.0 * batch * oh * oh * cout * 9 * cin
        return {"kernel": "conv3x3 %d->%d s%d @%dx%d%s" % (cin, cout, stride, h, h, " +BN stats" if stats else ""),
                "flops": flops, "ms": us * 1e-3, "tflops": flops / (us * 1e-6) / 1e12}

    parts = [  # (launches per step, measurement)
        (33, fwd(24, 64, 64, True)), (33, time_conv_dgrad(dev, batch, 24, 64, 64)),
        (2, fwd(96, 64, 64, False)), (1, time_conv_dgrad(dev, batch, 96, 64, 64)),
        (2, fwd(48, 64, 128, False)), (1, time_conv_dgrad(dev, batch, 48, 64, 128)),
        (2, fwd(48, 128, 128, False)), (1, time_conv_dgrad(dev, batch, 48, 128, 128)),
        (1, time_conv_dgrad(dev, batch, 24, 128, 256)),
        (3, fwd(48, 64, 128, True)), (3, time_conv_dgrad(dev, batch, 48, 64, 128)),
        (3, time_conv_dgrad(dev, batch, 24, 128, 256)),
        (3, fwd(96, 64, 64, True, 2)), (3, fwd(48, 128, 128, True, 2)),
        (2, time_conv_dgrad(dev, batch, 48, 64, 256, 1, 2)),
    ]
    if frozen:       # frozen trunk (configs[3]): no data gradient below the suffix conv
        parts = [pm for pm in parts if not pm[1]["kernel"].startswith("dgrad of conv3x3 64->64 s1 @24")]
    flops = sum(c * m["flops"] for c, m in parts)
    ms = sum(c * m["ms"] for c, m in parts)
    launches = sum(c for c, _ in parts)
    return {"kernel": "all %d launches per step (15 shape/direction classes, launch-count weighted)" % launches,
            "flops": flops / launches, "ms": ms / launches, "tflops": flops / (ms * 1e-3) / 1e12,
            "classes": [{"launches": c, "kernel": m["kernel"], "us": m["ms"] * 1e3, "tflops": m["tflops"]}
                        for c, m in parts]}


def time_hbm_kernels(dev, batch):
    """Achieved algorithmic GB/s of the fused BN / PReLU kernels on the largest activation of the step
    (generator up-scale output, batch x 96 x 96 x 64 bf16 = 75 MB per tensor)."""
    import torch
    from sisr_b200 import _lib
    rows, c = batch * 96 * 96, 64
    sets = [(torch.randn(rows, c, device=dev).to(torch.bfloat16), torch.randn(rows, c, device=dev).to(torch.bfloat16),
             torch.empty(rows, c, device=dev, dtype=torch.bfloat16)) for _ in range(3)]
    scale, shift = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
    mean, invstd = torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5
    slope = torch.full((1,), 0.25, device=dev)
    sums, colsum = torch.zeros(2 * c + 1, device=dev), torch.zeros(c, device=dev)
    nb = rows * c * 2
    cases = [
        ("bn_apply+PReLU", 2 * nb, lambda i, s: _lib.call("sisr_bn_apply", sets[i % 3][0], scale, shift, 3, 0.0, slope,
                                                           None, sets[i % 3][2], rows, c, s)),
        ("bn_apply+residual", 3 * nb, lambda i, s: _lib.call("sisr_bn_apply", sets[i % 3][0], scale, shift, 0, 0.0, None,
                                                              sets[i % 3][1], sets[i % 3][2], rows, c, s)),
        ("bn_bwd_reduce", 2 * nb, lambda i, s: _lib.call("sisr_bn_bwd_reduce", sets[i % 3][1], sets[i % 3][0], mean, invstd,
                                                          scale, shift, 3, 0.0, slope, sums, rows, c, s)),
        ("bn_bwd_apply", 3 * nb, lambda i, s: _lib.call("sisr_bn_bwd_apply", sets[i % 3][1], sets[i % 3][0], mean, invstd,
                                                         scale, shift, 3, 0.0, slope, sums, float(rows), sets[i % 3][2],
                                                         colsum, rows, c, s)),
        ("act_bwd(PReLU)", 3 * nb, lambda i, s: _lib.call("sisr_act_bwd", sets[i % 3][1], sets[i % 3][0], 3, 0.0, slope,
                                                           sets[i % 3][2], None, colsum, rows, c, s)),
    ]
    out = []
    for name, nbytes, fn in cases:
        us = _graph_time_us(fn, reps=6, replays=3)
        out.append({"kernel": name, "bytes": nbytes, "us": us, "GBps": nbytes / (us * 1e-6) / 1e9})
    return out


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel's most frequent class
    (trunk conv), parsed from the newest committed `ncu --set full` summary under profiles/ (null if none)."""
    import glob
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_trunk_conv*.txt")), reverse=True):
        rd = wr = None
        if "igemm_pm_kernel" not in open(path).read():
            # a capture of a kernel the trunk no longer runs (round 1: igemm_t_kernel) says nothing about this one;
            # round 2 ended without an `ncu --set full` capture of igemm_pm_kernel (GPU budget), hence null.  What the
            # older captures showed for the same tensors: 4.8 MB of DRAM traffic per launch against 9.5 MB algorithmic
            # (the output stays in L2) - profiles/r1_ncu_trunk_conv.txt, profiles/r2_ncu_partial_summary.txt
            continue
        for ln in open(path):
            m = re.search(r"dram__bytes_(read|write)\.sum\s+([0-9.,]+)\s+(\w+)", ln)
            if m:
                val = float(m.group(2).replace(",", "")) * unit.get(m.group(3), 1.0)
                if m.group(1) == "read" and rd is None:
                    rd = val
                elif m.group(1) == "write" and wr is None:
                    wr = val
            if rd is not None and wr is not None:
                return {"bytes": int(rd + wr), "source": os.path.relpath(path, ROOT)}
    return {"bytes": None, "source": None}


def dominant_share():
    """The dominant kernels' share of the step's serialised kernel time, read from the newest committed ncu
    launch-list summary (tools/launch_summary.py over `ncu --metrics gpu__time_duration.sum`)."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_launches_b64*summary.txt")), reverse=True)
    for path in paths:
        try:
            hits = [ln.split() for ln in open(path) if re_kernel(ln)]
        except OSError:
            continue
        if hits:
            parts = [f"{f[-1]} = {f[2]} of the step's kernel time, {f[3]} launches" for f in hits]
            return "; ".join(parts) + f" (ncu launch list, {os.path.relpath(path, ROOT)})"
    return "see profiles/"


def re_kernel(line):
    return any(k in line for k in ("igemm_t_kernel", "igemm_th_kernel", "igemm_pm_kernel"))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sisr_b200 as m  # noqa: F401
    from sisr_b200 import _lib, parallel

    rank, local, world = parallel.init_distributed()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~1 s to start; samples are windowed later
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = CONFIGS[args.config]
    batch = args.batch or cfg["batch"]
    HR, LRS = cfg["hr"], cfg["lr"]
    use_graph = not args.no_graph      # NCCL all-reduces and the peer-memory SyncBN kernels are captured too
    bucket = int(os.environ.get("SISR_BUCKET_MB", "0")) << 20
    gs = (parallel.GradSync(bucket_bytes=bucket) if bucket else parallel.GradSync()) if world > 1 else None
    tr = build_trainer(dev, batch, world, gs, args.config)
    gen = torch.Generator().manual_seed(1234 + rank)          # synthetic HR patches ~ U[-1, 1]
    hr_host = (torch.rand((batch, 3, HR, HR), generator=gen) * 2 - 1).pin_memory()
    import torch.nn.functional as F
    lr_host = F.interpolate(hr_host, (LRS, LRS), mode="bicubic", align_corners=True).clamp(-1, 1).pin_memory()
    hr, lr = hr_host.to(dev), lr_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    calls_before = _lib.LAUNCHES[0] if hasattr(_lib, "LAUNCHES") else 0
    if use_graph:
        tr.capture(hr, lr, warmup=2)
        step = lambda: tr.replay(hr, lr)         # noqa: E731
    else:
        step = lambda: tr.step(hr, lr)           # noqa: E731
    launches_per_step = None
    if hasattr(_lib, "LAUNCHES"):
        c0 = _lib.LAUNCHES[0]
        if use_graph:
            launches_per_step = (c0 - calls_before) // 3    # 2 warm-up steps + 1 captured step
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    # ---- timed region 1: device-resident inputs (inputs > L2? no: 126 MB L2 is flushed by the step
    # itself: one step touches > 1 GB of activations and weights)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not use_graph and hasattr(_lib, "LAUNCHES"):
        c0 = _lib.LAUNCHES[0]
    barrier()
    if os.environ.get("SISR_PROFILE_TIMED_REGION"):      # ncu --profile-from-start off
        torch.cuda.profiler.start()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    if os.environ.get("SISR_PROFILE_TIMED_REGION"):
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    if not use_graph and hasattr(_lib, "LAUNCHES"):
        launches_per_step = (_lib.LAUNCHES[0] - c0) // args.steps
    t = torch.tensor([ms], device=dev)
    ms_by_rank = [ms / args.steps]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        ms_by_rank = [float(x) / args.steps for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    # ---- timed region 2: end to end from pinned host buffers through the public API
    # (HostFeed + SRGANTrainer.replay_from_feed): every step copies its HR batch host -> device inside the
    # region (on the feed's copy stream, overlapping the previous step), makes LR on the device as
    # train.py:46 does, runs the step and reads the three losses back to the host (one step late, so the
    # host never stalls the queue; the last read is inside the region as well).
    from sisr_b200 import lr_from_hr
    from sisr_b200.train import HostFeed

    def loss_vec(o):
        return torch.stack([o["err_d"].reshape(()), o["err_g_adv"].reshape(()), o["err_g_cont"].reshape(())])

    feed = HostFeed(tuple(hr_host.shape), dev) if use_graph else None      # buffers allocated outside the timed region
    pin = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def e2e_pipelined():
        losses = None
        feed.submit(hr_host)
        for i in range(args.steps):
            if i + 1 < args.steps:
                feed.submit(hr_host)
            o = tr.replay_from_feed(feed)
            pin[i % 2].copy_(loss_vec(o), non_blocking=True)              # D2H
            done[i % 2].record()
            if i:
                done[(i - 1) % 2].synchronize()
                losses = pin[(i - 1) % 2].clone()
        done[(args.steps - 1) % 2].synchronize()
        return pin[(args.steps - 1) % 2].clone()

    def e2e_serial():
        losses = None
        for _ in range(args.steps):
            hr.copy_(hr_host, non_blocking=True)
            lr.copy_(lr_from_hr(hr, (LRS, LRS)))
            o = step()
            losses = loss_vec(o).cpu()                                     # D2H + sync
        return losses

    def timed_e2e(fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ls = fn()
        b.record()
        barrier()
        t_ = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_), ls

    ms_e2e_serial, losses = timed_e2e(e2e_serial)
    ms_e2e = ms_e2e_serial
    e2e_mode = "serial"
    ms_pipe = None
    if use_graph:
        ms_pipe, losses = timed_e2e(e2e_pipelined)
        if ms_pipe < ms_e2e:
            ms_e2e, e2e_mode = ms_pipe, "pipelined (HostFeed: H2D of batch i+1 overlaps step i)"
    clocks = sampler.window(wall0, time.time()) if rank == 0 else None
    # ---- sustained leg: the driver-dictated K steps last ~0.2 s (burst clocks); the same loop for >= S seconds
    # shows what the step costs once the part sits at its power cap
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds / (ms / args.steps * 1e-3)) + 1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ws0 = time.time()
        s0.record()
        for _ in range(n_sus):
            step()
        s1.record()
        barrier()
        ws1 = time.time()
        t_ = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms_sus = float(t_) / n_sus
        sustained = {"seconds": float(t_) * 1e-3, "steps": n_sus, "ms_per_step": ms_sus,
                     "value": batch * world / (ms_sus * 1e-3), "unit": "patches/s",
                     "clocks": sampler.window(ws0 + 1.0, ws1) if rank == 0 else None}
    # ---- the graph path with replayed fakes (train.py:144-146: after 100 iterations int(len * 0.01) old fake
    # batches join every D update; one CUDA graph per count k, captured on first use)
    replay_path = None
    if world == 1 and use_graph and args.config == "x4" and not args.no_replay_path:
        replay_path = {}
        for k in (1, 10):
            olds = [(torch.rand((batch, 3, HR, HR), device=dev) * 2 - 1).to(torch.bfloat16) for _ in range(k)]
            tr.replay(hr, lr, old_fakes=olds)                    # captures the k-fake graph, then replays it
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(10):
                tr.replay(hr, lr, old_fakes=olds)
            r1.record()
            barrier()
            replay_path[f"k={k}"] = {"ms_per_step": r0.elapsed_time(r1) / 10,
                                     "d_passes_per_step": 3 + k}
            del olds
    if rank == 0:
        sampler.stop()
    dp = dp_check(dev, rank, world, tr) if (world > 1 or os.environ.get("SISR_WRITE_DP_GOLDEN")) else None
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    peaks = measured_peaks()
    traffic = ncu_traffic()
    if args.config == "x8":
        # the FLOP-heaviest conv class of the 256x256 step (VGG conv3_x: 256 -> 256 @64x64, 10 launches per step)
        dom = time_conv_kernel(dev, batch, 64, 256, 256, False)
        dom["classes"] = [{"launches": 10, "kernel": dom["kernel"], "us": dom["ms"] * 1e3, "tflops": dom["tflops"]}]
        dom["kernel"] = "igemm_tc_kernel<256,4>: " + dom["kernel"]
        traffic = {"bytes": None, "source": None}
        others = [time_conv_kernel(dev, batch, 256, 64, 64, False), time_conv_kernel(dev, batch, 128, 128, 128, False),
                  time_conv_kernel(dev, batch, 32, 512, 512, False)]
        hbm = []
    else:
        dom = time_dominant_kernel(dev, batch, frozen=cfg["frozen"])
        dom["kernel"] = "igemm_t_kernel / igemm_pm_kernel: " + dom["kernel"]
        others = [time_conv_kernel(dev, batch, 96, 64, 64, False), time_conv_kernel(dev, batch, 48, 128, 128, False),
                  time_conv_kernel(dev, batch, 24, 256, 256, False), time_conv_kernel(dev, batch, 12, 512, 512, False)]
        hbm = time_hbm_kernels(dev, batch)
    total_patches = batch * world * args.steps
    value = total_patches / (ms * 1e-3)
    line = {
        "metric": "SRGAN x4 train-step HR patches/sec", "value": value, "unit": "patches/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "ms_per_step_by_rank": ms_by_rank,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config,
                   "batch_per_gpu": batch, "global_batch": batch * world, "spectral_norm": True,
                   "parallelism": f"dp{world}" + (" + SyncBN" if world > 1 else ""),
                   "cuda_graph": use_graph,
                   "l2": "one step streams > 1 GB of activations/weights (> 126 MB L2) between reuses"},
        "flop_per_patch": cfg["flop"],
        "step_tflops": cfg["flop"] * batch / (ms / args.steps * 1e-3) / 1e12,
        "step_frac_of_sustained_bf16": cfg["flop"] * batch / (ms / args.steps * 1e-3) / 1e12
        / peaks["bf16_sustained"],
        "sustained": sustained,
        "replay_path": replay_path,
        "losses": [float(x) for x in losses],
        "dp_check": dp,
        "clocks": clocks,
        "e2e": {"value": total_patches / (ms_e2e * 1e-3), "unit": "patches/s",
                "h2d_bytes_per_step": hr_host.numel() * 4,
                "d2h_bytes_per_step": 12, "mode": e2e_mode,
                "serial_value": total_patches / (ms_e2e_serial * 1e-3),
                "pipelined_value": (total_patches / (ms_pipe * 1e-3)) if ms_pipe else None},
        "gpu_launches": (launches_per_step or 0) * args.steps,
        "roofline": {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["bf16_burst"],
                     "unit": "TFLOP/s", "frac": dom["tflops"] / peaks["bf16_burst"],
                     # dram__bytes_read + write of one launch of the most frequent class (trunk conv), ncu
                     # --set full (profiles/r1_ncu_trunk_conv.txt): 4.83 MB read (the input activation,
                     # cold), 0 written back (the output stays in L2)
                     "traffic": traffic["bytes"], "traffic_source": traffic["source"],
                     "kernel": dom["kernel"], "ms_per_launch": dom["ms"],
                     "flops_per_launch": dom["flops"], "classes": dom["classes"],
                     "peak_source": peaks["source"] + " (burst: kernel timed alone)",
                     "share_of_step": dominant_share(),
                     "timing": "12 back-to-back launches per CUDA graph over 3 rotating buffer sets (tensors "
                               "L2-resident as in the step), CUDA events on the replaying stream"},
        "roofline_other": [{"kernel": o["kernel"], "achieved": o["tflops"], "unit": "TFLOP/s",
                            "frac": o["tflops"] / peaks["bf16_burst"], "ms_per_launch": o["ms"]} for o in others],
        "roofline_hbm": [{"kernel": h["kernel"] + " on batch x 96x96x64 bf16", "bound": "hbm",
                          "achieved": h["GBps"], "peak": peaks["hbm"], "unit": "GB/s",
                          "frac": h["GBps"] / peaks["hbm"], "us_per_launch": h["us"]} for h in hbm],
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        got = reference_train_loop_time(16, 2, 1, threads, budget_s=30.0)
        if got is not None:
            line["cpu_baseline"] = {"value": 16 / got[0], "unit": "patches/s", "cores": threads, "kind": "reference",
                                    "sample": f"{got[1]} steps of batch 16 after 1 warm-up step: the unmodified "
                                              "reference train.train_loop (oracle/_ref), fp32, all host threads; "
                                              "`--impl reference` runs it at batch 64"}
        else:
            sec = cpu_reference_step_time(8, 1, 1, threads)
            line["cpu_baseline"] = {"value": 8 / sec, "unit": "patches/s", "cores": threads, "kind": "port",
                                    "sample": "1 step of batch 8 after 1 warm-up step (oracle port of the "
                                              "reference step, fp32, all host threads; oracle/_ref absent)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        # no destroy_process_group(): the captured graph still references the NCCL communicator; the
        # "destroy_process_group() was not called" warning on stderr at exit is harmless
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="patches per GPU (default: the config's)")
    ap.add_argument("--config", default="x4", choices=sorted(CONFIGS),
                    help="x4 = BASELINE.json configs[1]/[2]; frozen = configs[3]; x8 = configs[4]")
    ap.add_argument("--sustained-seconds", type=float, default=5.0)
    ap.add_argument("--no-replay-path", action="store_true",
                    help="skip timing the CUDA-graph step with 1 and 10 replayed fakes")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
