"""Host logic of the fused Adam's checkpoint interchange with torch.optim.Adam (+ LambdaLR), both
directions, on the CPU: no kernel runs here (the CUDA update itself is covered by tests/test_gpu_ops.py
and the stepping-after-cross-load case by tests/test_gpu_models.py).
Reference: config.py:170-180 (LambdaLR rebuilt on resume), 293-302 (Adam + best-effort state load),
utils.py:107-112 (what is saved)."""
import copy

import torch


def _params():
    g = torch.Generator().manual_seed(3)
    return [torch.nn.Parameter(torch.randn(4, 3, generator=g)), torch.nn.Parameter(torch.randn(5, generator=g))]


def test_reference_adam_steps_after_loading_our_state():
    import sisr_b200 as m
    ours = m.Adam(_params(), lr=1e-5, betas=(0.9, 0.999), decay_per_step=0.999)
    for p in ours.param_groups[0]["params"]:       # state as the CUDA step would have left it
        ours.state[p] = {"exp_avg": torch.full_like(p, 0.1), "exp_avg_sq": torch.full_like(p, 0.01)}
    sd = copy.deepcopy(ours.state_dict())
    group = sd["param_groups"][0]
    for k in ("weight_decay", "amsgrad", "maximize", "foreach", "capturable", "differentiable", "fused"):
        assert k in group, k
    ref_params = _params()
    ref = torch.optim.Adam(ref_params, lr=1e-5, betas=(0.9, 0.999))
    ref.load_state_dict(sd)
    sched = torch.optim.lr_scheduler.LambdaLR(ref, lr_lambda=lambda it: 0.5 ** it)
    before = [p.detach().clone() for p in ref_params]
    for p in ref_params:
        p.grad = torch.ones_like(p)
    ref.step()                                     # raised KeyError('weight_decay') before the fix
    sched.step()
    assert all(not torch.equal(a, b.detach()) for a, b in zip(before, ref_params))
    assert abs(ref.param_groups[0]["initial_lr"] - 1e-5) < 1e-12


def test_our_adam_accepts_a_reference_state_and_restarts_the_schedule():
    import sisr_b200 as m
    ref_params = _params()
    ref = torch.optim.Adam(ref_params, lr=1e-5, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.LambdaLR(ref, lr_lambda=lambda it: 0.5 ** it)
    for _ in range(3):
        for p in ref_params:
            p.grad = torch.ones_like(p)
        ref.step()
        sched.step()
    sd = ref.state_dict()
    assert abs(sd["param_groups"][0]["lr"] - 1e-5 * 0.125) < 1e-12      # the decayed value is what torch saves
    ours = m.Adam(_params(), lr=1e-5, betas=(0.9, 0.999), decay_per_step=0.999)
    ours.load_state_dict(sd)
    g = ours.param_groups[0]
    assert g["decay_per_step"] == 0.999                                  # KeyError at step() before the fix
    assert abs(g["lr"] - 1e-5) < 1e-12                                   # restart from initial_lr, as config.py does
    st = ours.state[g["params"][0]]
    assert int(st["step"]) == 3 and torch.allclose(st["exp_avg"], ref.state[ref_params[0]]["exp_avg"])
    # and our own round trip keeps everything
    again = m.Adam(_params(), lr=3e-4, decay_per_step=1.0)
    again.load_state_dict(ours.state_dict())
    assert abs(again.param_groups[0]["lr"] - 1e-5) < 1e-12 and again.param_groups[0]["decay_per_step"] == 0.999
