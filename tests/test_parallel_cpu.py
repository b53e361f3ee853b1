"""CPU, world_size 2, gloo: the bucketed gradient all-reduce averages gradients across ranks and
hands flat views to the optimizer; SyncBN statistic reduction goes through the same group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sisr_b200  # noqa: F401
    from sisr_b200 import ops, parallel
    r, _, w = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    if rank == 1:                                     # replicas start different ...
        with torch.no_grad():
            for p_ in net.parameters():
                p_.add_(1.0)

    class FakeOpt:
        def __init__(self, params):
            self.param_groups = [{"params": list(params)}]
            self.grad_views, self.grad_scale = None, 1.0
    opt = FakeOpt(net.parameters())
    sync = parallel.GradSync(bucket_bytes=300)       # several buckets
    sync.attach(opt, net)                            # ... attach(module=...) broadcasts rank 0's weights
    sync.attach(opt, net)                            # idempotent: no second set of hooks
    w0 = [p_.detach().clone() for p_ in net.parameters()]
    same = [[torch.zeros_like(t) for _ in range(world)] for t in w0]
    for t, out in zip(w0, same):
        dist.all_gather(out, t)
    assert all(torch.equal(o[0], o[1]) for o in same)
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 8)
    net(x).pow(2).sum().backward()
    sync.sync(opt)
    local = [p.grad.clone() for p in net.parameters()]
    summed = [opt.grad_views[p].clone() for p in net.parameters()]
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, out in zip(local, gathered):
        dist.all_gather(out, g)
    ok = all(torch.allclose(s, sum(o), atol=1e-6) for s, o in zip(summed, gathered))
    ok = ok and abs(opt.grad_scale - 1.0 / world) < 1e-12
    # SyncBN reduction helper
    t = torch.full((4,), float(rank + 1))
    ops._all_reduce(t)
    ok = ok and torch.equal(t, torch.full((4,), 3.0))
    # SyncBN is opt-in per step: outside the trainer's scope a BN forward uses local statistics
    ok = ok and ops._world() == 1
    with ops.sync_bn_scope():
        ok = ok and ops._world() == world
    # gradients that bypass autograd reach the buckets through the public callback
    ok = ok and ops._grad_ready_cb[0] is not None and ops._grad_ready_cb[0].__self__ is sync
    # end-of-backward hook: GradSync.drain is registered, only runs when side-stream gradients are handed over
    # (begin_step() calls join_wgrad() too - as the first thing of a CUDA-graph capture), and only waits for
    # all-reduces launched since the last sync()
    ok = ok and ops._before_join_cb[0] is not None and ops._before_join_cb[0].__self__ is sync
    ok = ok and sync._in_flight is False
    calls = []
    ops.set_before_join_callback(lambda: calls.append(1))
    ops.join_wgrad()
    ok = ok and calls == []
    ops._async["pending"]["k"] = (None, None, None, None, False, False)
    ops.join_wgrad()
    ok = ok and calls == [1] and not ops._async["pending"]
    ops.set_before_join_callback(sync.drain)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_sync_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
