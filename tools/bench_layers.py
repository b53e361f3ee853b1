"""Per-layer timing of the conv kernels of the SRGAN step through the C ABI (GPU).
Each shape is launched REP times inside one CUDA graph (no host gaps, tensor maps encoded at
capture), rotating over 3 buffer sets; reports microseconds per launch and TFLOP/s.
usage: python tools/bench_layers.py [fprop|dgrad|wgrad|all] [name-substring]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sisr_b200 import _lib

B = int(os.environ.get("BATCH", "64"))
LAYERS = [  # name, h, cin, cout, stride, ps_r, stats, count per step (fwd)
    ("G trunk 64->64 @24", 24, 64, 64, 1, 0, True, 33),
    ("G up 64->256 ps @24", 24, 64, 256, 1, 2, False, 1),
    ("G up 64->256 ps @48", 48, 64, 256, 1, 2, False, 1),
    ("D 64->64 s2 @96", 96, 64, 64, 2, 0, True, 3),
    ("D 64->128 @48", 48, 64, 128, 1, 0, True, 3),
    ("D 128->128 s2 @48", 48, 128, 128, 2, 0, True, 3),
    ("D 128->256 @24", 24, 128, 256, 1, 0, True, 3),
    ("D 256->256 s2 @24", 24, 256, 256, 2, 0, True, 3),
    ("D 256->512 @12", 12, 256, 512, 1, 0, True, 3),
    ("D 512->512 s2 @12", 12, 512, 512, 2, 0, True, 3),
    ("V 64->64 @96", 96, 64, 64, 1, 0, False, 2),
    ("V 64->128 @48", 48, 64, 128, 1, 0, False, 2),
    ("V 128->128 @48", 48, 128, 128, 1, 0, False, 2),
    ("V 128->256 @24", 24, 128, 256, 1, 0, False, 2),
    ("V 256->256 @24", 24, 256, 256, 1, 0, False, 6),
    ("V 256->512 @12", 12, 256, 512, 1, 0, False, 2),
    ("V 512->512 @12", 12, 512, 512, 1, 0, False, 6),
    ("V 512->512 @6", 6, 512, 512, 1, 0, False, 8),
    ("thin 3->64 @96", 96, 3, 64, 1, 0, False, 5),
    ("thin 64->3 @96", 96, 64, 3, 1, 0, False, 1),
]
REP = 12


def run(kind, name, h, cin, cout, stride, ps, stats):
    dev = torch.device("cuda")
    oh = (h + 2 - 3) // stride + 1
    d = _lib.ConvDesc(B, h, h, cin, oh, oh, cout, 3, stride, 1, ps)
    sets = []
    for _ in range(3):
        x = torch.randn(B, h, h, cin, device=dev).to(torch.bfloat16)
        wf = (torch.randn(cout, 3, 3, cin, device=dev) * 0.02).to(torch.bfloat16)
        wd = (torch.randn(cin, 3, 3, cout, device=dev) * 0.02).to(torch.bfloat16)
        if ps == 2:
            y = torch.randn(B, oh * 2, oh * 2, cout // 4, device=dev).to(torch.bfloat16)
        else:
            y = torch.randn(B, oh, oh, cout, device=dev).to(torch.bfloat16)
        dx = torch.empty_like(x)
        st = torch.empty(_lib.query("sisr_stats_rows") * 2 * cout, device=dev) if stats else None
        gp = torch.empty(cout, 3, 3, cin, device=dev)
        db = torch.empty(cout, device=dev)
        ws = torch.empty(max(_lib.query("sisr_conv_wgrad_workspace_bytes", d), 4), dtype=torch.uint8, device=dev)
        sets.append((x, wf, wd, y, dx, st, gp, db, ws))
    bias = torch.zeros(cout, device=dev)

    def launch(i, s):
        x, wf, wd, y, dx, st, gp, db, ws = sets[i % 3]
        if kind == "fprop":
            _lib.call("sisr_conv_fprop", d, x, wf, bias, 0, 0.0, None, y, None, st, s)
        elif kind == "dgrad":
            _lib.call("sisr_conv_dgrad", d, y, wf, wd, dx, s)
        else:
            _lib.call("sisr_conv_wgrad", d, x, y, gp, db, ws, s)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(3):
            launch(i, side.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for i in range(REP):
            launch(i, s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * REP)
    flops = 2.0 * B * oh * oh * cout * 9 * cin
    return us, flops / (us * 1e-6) / 1e12


def main():
    if os.environ.get("PM_MODE"):
        _lib.query("sisr_debug_pm_mode", int(os.environ["PM_MODE"]))
    if os.environ.get("TRANSPOSED"):
        _lib.query("sisr_debug_transposed", int(os.environ["TRANSPOSED"]))
    kinds = ["fprop", "dgrad", "wgrad"] if len(sys.argv) < 2 or sys.argv[1] == "all" else [sys.argv[1]]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    tot = {k: 0.0 for k in kinds}
    for (name, h, cin, cout, stride, ps, stats, cnt) in LAYERS:
        if filt not in name:
            continue
        row = f"{name:24s}"
        for k in kinds:
            us, tf = run(k, name, h, cin, cout, stride, ps, stats)
            tot[k] += us * cnt
            row += f"  {k} {us:8.1f} us {tf:7.1f} TF/s"
        print(row, flush=True)
    print("weighted by per-step forward count:", {k: round(v / 1e3, 3) for k, v in tot.items()}, "ms")


if __name__ == "__main__":
    main()
