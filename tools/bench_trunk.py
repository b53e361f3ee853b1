"""Generator forward (train mode, no autograd) with the trunk fused into one persistent kernel vs layer by layer:
CUDA-graph replays timed with CUDA events.  usage: python tools/bench_trunk.py [batch] [lr_size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sisr_b200 as m
from sisr_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda")
torch.manual_seed(0)
net = m.GeneratorSuffix(m.Generator(16, 64, 256, [2], use_sn=True)).to(dev).train()
x = (torch.rand(B, 3, S, S, device=dev) * 2 - 1)
out = {}
for fused in (False, True, False, True):
    ops.set_fused_trunk(fused)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):
            y = net(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = _lib.LAUNCHES[0]
    with torch.cuda.graph(g), torch.no_grad():
        y = net(x)
    launches = _lib.LAUNCHES[0] - n0
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    out.setdefault(fused, []).append(ms)
    print(f"fused={fused}: generator forward {ms * 1e3:.1f} us, {launches} launches", flush=True)
a, b = min(out[False]), min(out[True])
print(f"trunk fused saves {1e3 * (a - b):.1f} us per generator forward ({a * 1e3:.1f} -> {b * 1e3:.1f} us); "
      f"33 layers: {1e3 * (a - b) / 33:.2f} us per layer")
