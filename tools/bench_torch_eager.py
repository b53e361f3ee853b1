"""PyTorch-eager (cuDNN / cuBLAS) baseline of the SAME step on the SAME B200 - measurement infrastructure only
(SURVEY.md 2.1 names it as the bar to beat; nothing here is on the product path).

  1. ``reference-fp32``: the UNMODIFIED reference ``train.train_loop`` (oracle/_ref) with every module on
     ``cuda`` - PyTorch's defaults (cuDNN convolutions with TF32 allowed, fp32 storage), its per-step ``.item()``
     syncs and CPU replay list included, exactly what a user of the reference gets on this GPU;
  2. ``reference-fp32-no-tf32``: the same with ``torch.backends.cudnn.allow_tf32 = False`` (true fp32 math);
  3. ``oracle-bf16-autocast``: the oracle's functional restatement of the step under
     ``torch.autocast(bfloat16)`` with channels_last inputs (the reference's own loop cannot run under
     autocast: nn.BCELoss refuses half-precision inputs).

    python tools/bench_torch_eager.py [--batch 64] [--steps 10] [--warmup 3]

Prints one JSON line per variant: ms/step and HR patches/s, config 2 (G 16 blocks + suffix, D @96, VGG54).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

D_FEATS = [64, 64, 128, 128, 256, 256, 512, 512]
D_STRIDES = [1, 2, 1, 2, 1, 2, 1, 2]
VGG54 = 0b10000


class Timed:
    def __init__(self, batches):
        self.batches, self.stamps = batches, []

    def __iter__(self):
        import torch
        for b in self.batches:
            torch.cuda.synchronize()
            self.stamps.append(time.perf_counter())
            yield b

    def __len__(self):
        return len(self.batches)


def reference_loop(batch, steps, warmup, allow_tf32):
    import torch
    from oracle import ref_harness as R
    from oracle import state_factory as S
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    mg, md, mc = R.import_reference()
    net_g = mg.GeneratorSuffix(mg.Generator(16, 64, 256, [2], use_sn=True))
    net_d = md.Discriminator((3, 96, 96), D_FEATS, D_STRIDES)
    ext = mc.MaskedVGG(VGG54)
    torch.nn.Module.load_state_dict(net_g, S.generator_state(1, n_blocks=16, n_suffix=1), strict=True)
    torch.nn.Module.load_state_dict(net_d, S.discriminator_state(2, (3, 96, 96), D_FEATS, D_STRIDES), strict=True)
    torch.nn.Module.load_state_dict(ext, S.vgg_state(3, VGG54), strict=True)
    net_g, net_d, ext = net_g.cuda(), net_d.cuda(), ext.cuda()
    data = Timed([S.synthetic_hr(10 + i, batch, 96) for i in range(warmup + steps + 1)])
    R.run_train_loop(net_g, net_d, ext, data, lr=1e-5, lr_size=24, device="cuda")
    st = data.stamps
    return (st[warmup + steps] - st[warmup]) / steps


def oracle_autocast(batch, steps, warmup):
    import torch
    from oracle import srgan_oracle as O
    from oracle import state_factory as S
    dev = torch.device("cuda")
    g = {k: v.to(dev) for k, v in S.generator_state(1, n_blocks=16, n_suffix=1).items()}
    d = {k: v.to(dev) for k, v in S.discriminator_state(2, (3, 96, 96), D_FEATS, D_STRIDES).items()}
    v = {k: t.to(dev) for k, t in S.vgg_state(3, VGG54).items()}
    og, od = O.AdamState(O.trainable_names(g), 1e-5), O.AdamState(O.trainable_names(d), 1e-5)
    times = []
    for i in range(warmup + steps):
        hr = S.synthetic_hr(10 + i, batch, 96).to(dev).contiguous(memory_format=torch.channels_last)
        lr = O.lr_from_hr(hr, (24, 24)).contiguous(memory_format=torch.channels_last)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            O.train_step(g, d, v, hr, lr, d_strides=D_STRIDES, vgg_mask=VGG54, opt_g=og, opt_d=od)
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    name = torch.cuda.get_device_name(0)
    for variant, fn in (("reference-fp32 (cuDNN, TF32 allowed: PyTorch default)",
                         lambda: reference_loop(args.batch, args.steps, args.warmup, True)),
                        ("reference-fp32-no-tf32", lambda: reference_loop(args.batch, args.steps, args.warmup, False)),
                        ("oracle-bf16-autocast-channels_last", lambda: oracle_autocast(args.batch, args.steps, args.warmup))):
        try:
            sec = fn()
            print(json.dumps({"impl": "torch-eager", "variant": variant, "gpu": name, "batch": args.batch,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                              "value": args.batch / sec, "unit": "patches/s",
                              "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)
        except Exception as e:  # noqa: BLE001 - a variant that cannot run is reported, not fatal
            print(json.dumps({"impl": "torch-eager", "variant": variant, "error": repr(e)[:300]}), flush=True)


if __name__ == "__main__":
    main()
