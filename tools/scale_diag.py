"""Where does the data-parallel overhead go?  Times the captured config-2 step (64 patches per GPU) at the
launched world size under diagnostic switches that remove ONE kind of exchange each (the results of those runs
are not training steps - the tool only reads their duration):

  DIAG=base        the product path
  DIAG=nosyncbn    BatchNorm statistics stay local (no peer exchange), gradients still all-reduced
  DIAG=noreduce    SyncBN as in the product, gradient buckets packed but not all-reduced
  DIAG=neither     both removed (= a single-GPU step plus the bucket packing)

usage: DIAG=... python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/scale_diag.py [steps]
The switches are monkey patches applied here; the package has no such knobs."""
import contextlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

import bench
from sisr_b200 import ops, parallel


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    diag = os.environ.get("DIAG", "base")
    rank, local, world = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if diag in ("nosyncbn", "neither"):
        ops.sync_bn_scope = contextlib.nullcontext            # trainer.step looks it up through the module
    if diag in ("noreduce", "neither"):
        real = dist.all_reduce

        def fake_all_reduce(t, *a, **k):
            if t.numel() > 4096:                              # the gradient buckets; barriers / scalars go through
                return None
            return real(t, *a, **k)
        parallel.dist.all_reduce = fake_all_reduce
    bucket = int(os.environ.get("SISR_BUCKET_MB", "0")) << 20
    gs = (parallel.GradSync(bucket_bytes=bucket) if bucket else parallel.GradSync()) if world > 1 else None
    tr = bench.build_trainer(dev, 64, world, gs, "x4")
    gen = torch.Generator().manual_seed(1234 + rank)
    hr = (torch.rand((64, 3, 96, 96), generator=gen) * 2 - 1).to(dev)
    lr = F.interpolate(hr, (24, 24), mode="bicubic", align_corners=True).clamp(-1, 1)
    tr.capture(hr, lr, warmup=2)
    for _ in range(5):
        tr.replay(hr, lr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.replay(hr, lr)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        real_ar = dist.all_reduce if diag not in ("noreduce", "neither") else real
        real_ar(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"diag": diag, "world": world, "ms_per_step": round(float(t), 4),
                          "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")},
                          "bucket_mb": bucket >> 20}), flush=True)
    if world > 1:
        dist.barrier()
    # no destroy_process_group(): the captured graph still references the NCCL communicator and the teardown
    # then blocks until the launcher's timeout (that cost this tool's first run 150 s per variant)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
