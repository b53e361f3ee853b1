"""Timing of the HBM/L2-bound fused kernels (BN apply / backward, activation backward) through the
C ABI (GPU): REP launches per CUDA graph, rotating 3 buffer sets; reports us per launch and the
achieved algorithmic GB/s (bf16 elements read + written)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sisr_b200 import _lib

REP = 12
B = int(os.environ.get("BATCH", "64"))
SHAPES = [("G trunk 24x24x64", B * 24 * 24, 64), ("D conv1 48x48x64", B * 48 * 48, 64),
          ("D conv3 24x24x128", B * 24 * 24, 128), ("G up 96x96x64", B * 96 * 96, 64)]


def timed(fn):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(3):
            fn(i, side.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for i in range(REP):
            fn(i, s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * REP)


def run():
    dev = torch.device("cuda")
    out = []
    for name, rows, c in SHAPES:
        sets = []
        for _ in range(3):
            y = torch.randn(rows, c, device=dev).to(torch.bfloat16)
            g = torch.randn(rows, c, device=dev).to(torch.bfloat16)
            o = torch.empty_like(y)
            sets.append((y, g, o))
        scale = torch.rand(c, device=dev) + 0.5
        shift = torch.randn(c, device=dev)
        mean = torch.randn(c, device=dev) * 0.1
        invstd = torch.rand(c, device=dev) + 0.5
        slope = torch.full((1,), 0.25, device=dev)
        sums = torch.zeros(2 * c + 1, device=dev)
        colsum = torch.zeros(c, device=dev)
        nbytes = rows * c * 2
        cases = {
            "bn_apply+prelu": (lambda i, s: _lib.call("sisr_bn_apply", sets[i % 3][0], scale, shift, 3, 0.0, slope,
                                                      None, sets[i % 3][2], rows, c, s), 2 * nbytes),
            "bn_apply+residual": (lambda i, s: _lib.call("sisr_bn_apply", sets[i % 3][0], scale, shift, 0, 0.0, None,
                                                         sets[i % 3][1], sets[i % 3][2], rows, c, s), 3 * nbytes),
            "bn_bwd_reduce": (lambda i, s: _lib.call("sisr_bn_bwd_reduce", sets[i % 3][1], sets[i % 3][0], mean, invstd,
                                                     scale, shift, 3, 0.0, slope, sums, rows, c, s), 2 * nbytes),
            "bn_bwd_apply": (lambda i, s: _lib.call("sisr_bn_bwd_apply", sets[i % 3][1], sets[i % 3][0], mean, invstd,
                                                    scale, shift, 3, 0.0, slope, sums, float(rows), sets[i % 3][2],
                                                    colsum, rows, c, s), 3 * nbytes),
            "act_bwd": (lambda i, s: _lib.call("sisr_act_bwd", sets[i % 3][1], sets[i % 3][0], 3, 0.0, slope,
                                               sets[i % 3][2], None, colsum, rows, c, s), 3 * nbytes),
        }
        for k, (fn, bytes_) in cases.items():
            us = timed(fn)
            out.append({"kernel": k, "tensor": name, "us": us, "GBps": bytes_ / (us * 1e-6) / 1e9,
                        "MB": bytes_ / 1e6})
    return out


if __name__ == "__main__":
    for r in run():
        print(f"{r['tensor']:20s} {r['kernel']:18s} {r['us']:8.1f} us  {r['MB']:7.1f} MB  {r['GBps']:8.0f} GB/s", flush=True)
