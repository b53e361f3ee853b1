"""A-B check of the weight-gradient kernel's two accumulation modes on the trunk shape (run twice:
SISR_WGRAD_ATOMIC=0 writes the reference file, the default run compares against it).
usage: SISR_WGRAD_ATOMIC=0 python tools/check_wgrad_atomic.py /tmp/w.pt; python tools/check_wgrad_atomic.py /tmp/w.pt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sisr_b200 import _lib

path = sys.argv[1]
dev = torch.device("cuda")
torch.manual_seed(0)
res = {}
for (n, h, cin, cout) in ((64, 24, 64, 64), (8, 48, 64, 128), (5, 12, 128, 128)):
    d = _lib.ConvDesc(n, h, h, cin, h, h, cout, 3, 1, 1, 0)
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, h, h, cout, device=dev).to(torch.bfloat16)
    gp = torch.empty(cout, 3, 3, cin, device=dev)
    db = torch.empty(cout, device=dev)
    ws = torch.empty(max(_lib.query("sisr_conv_wgrad_workspace_bytes", d), 4), dtype=torch.uint8, device=dev)
    outs = []
    for rep in range(3):
        _lib.call("sisr_conv_wgrad", d, x, dy, gp, db, ws, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append(gp.clone().cpu())
    res[(n, h, cin, cout)] = outs
mode = "ordered" if os.environ.get("SISR_WGRAD_ATOMIC") == "0" else "atomic"
if mode == "ordered":
    torch.save(res, path)
    print("saved ordered-mode gradients")
else:
    ref = torch.load(path)
    for k, outs in res.items():
        r = ref[k][0].double()
        rel = [float((o.double() - r).norm() / r.norm()) for o in outs]
        rr = float((outs[0].double() - outs[1].double()).norm() / r.norm())
        print(f"shape {k}: atomic vs ordered rel-L2 {max(rel):.2e}; atomic run-to-run {rr:.2e}; "
              f"ordered run-to-run {float((ref[k][0].double() - ref[k][1].double()).norm() / r.norm()):.2e}")
        assert max(rel) < 1e-5, rel
    print("WGRAD_ATOMIC_CHECK PASS")
