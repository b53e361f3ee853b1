"""Where does the tensor-core weight-gradient kernel spend its cycles?  Phase counters of CTA 0 (producer waits for
a free stage / MMA warp waits for operands / MMA warp issues / epilogue) and the launch time, per layer shape.
usage: python tools/wgrad_phases.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sisr_b200 import _lib

dev = torch.device("cuda")
torch.manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
SHAPES = [("G trunk 64->64 @24", 64, 24, 64, 64, 1), ("D 64->128 @48", 64, 48, 64, 128, 1),
          ("D 128->128 s2 @48", 64, 48, 128, 128, 2), ("D 128->256 @24", 64, 24, 128, 256, 1),
          ("D 256->512 @12", 64, 12, 256, 512, 1), ("D 512->512 s2 @12", 64, 12, 512, 512, 2)]
for name, n, h, cin, cout, stride in SHAPES:
    oh = (h + 2 - 3) // stride + 1
    d = _lib.ConvDesc(n, h, h, cin, oh, oh, cout, 3, stride, 1, 0)
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, oh, oh, cout, device=dev).to(torch.bfloat16)
    gp = torch.empty(cout, 3, 3, cin, device=dev)
    db = torch.empty(cout, device=dev)
    ws = torch.empty(max(_lib.query("sisr_conv_wgrad_workspace_bytes", d), 4), dtype=torch.uint8, device=dev)
    run = lambda: _lib.call("sisr_conv_wgrad", d, x, dy, gp, db, ws, st)   # noqa: E731
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    cnt = torch.zeros(8, dtype=torch.int64, device=dev)
    _lib.query("sisr_debug_wgrad_counters", cnt)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    _lib.query("sisr_debug_wgrad_counters", None)
    c = cnt.tolist()
    kb, launches = max(c[3], 1), max(c[6], 1)
    flops = 2.0 * n * oh * oh * cout * 9 * cin
    print(f"{name:22s} {us:7.1f} us/launch (wgrad + split reduce + bias) {flops / us / 1e6:7.1f} TFLOP/s | grid {c[7]}, "
          f"{kb // launches} k-blocks per CTA; per k-block: producer waits {c[0] // kb}, MMA warp waits {c[1] // kb}, "
          f"issues {c[2] // kb} | epilogue {c[4] // launches}, kernel {c[5] // launches} cycles", flush=True)
