"""Stream schedule of the step: runs last (file name) - added at the very end of round 2, after the
GPU budget was spent, so a surprise here must not hide the results of the other files under `pytest -x`."""
import pytest
import torch

from oracle import state_factory as S
from test_gpu_models import _build_step, _load, rel

pytestmark = pytest.mark.gpu


def test_step_schedule_streams_do_not_change_the_step(cuda, golden_dir):
    """D(real) and MaskedVGG(fake) on their own streams (StepConfig.overlap_d_real / overlap_fake_features,
    DESIGN.md 4a) against the same step on one stream (+ the weight-gradient stream): same losses for three steps.
    lr = 1e-5 as in test_host_feed_pipeline_matches_direct_replay (two runs differ in the last bits of a few fp32
    reductions; Adam's first updates are sign-like)."""
    import sisr_b200 as m
    g = _load(golden_dir, "train_step2")
    hrs = [S.synthetic_hr(g["seed"] + 60 + i, g["B"], g["HR"]) for i in range(3)]
    keys = ("err_d", "err_g_adv", "err_g_cont")
    res = []
    for overlap in (False, True):
        tr, _ = _build_step(m, g["seed"], g["shape"], g["features"], g["strides"], g["mask"], 1e-5)
        tr.cfg.overlap_d_real = tr.cfg.overlap_fake_features = overlap
        outs = []
        for h in hrs:
            hd = h.cuda()
            o = tr.step(hd, m.lr_from_hr(hd, (g["LR"], g["LR"])))
            outs.append([float(o[k]) for k in keys])
        torch.cuda.synchronize()
        stats = {k: v.detach().float().cpu().clone() for k, v in tr.net_d.state_dict().items() if "running" in k}
        res.append((outs, stats))
    for a, b in zip(res[0][0], res[1][0]):
        for x, y in zip(a, b):
            assert abs(x - y) <= 1e-2 * abs(x) + 1e-6, (res[0][0], res[1][0])
    # the discriminator's BN running statistics saw the same three passes per step in the same order
    for k, v in res[0][1].items():
        assert rel(res[1][1][k], v) < 1e-3, k
