"""Run the UNMODIFIED reference step (train.train_loop) on synthetic data.  TEST INFRASTRUCTURE ONLY.

The reference is eight flat Python scripts with no build system.  ``make_ref()`` is the committed recipe
that places a copy of them under ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the GPU
box with the snapshot, /root/reference itself does not exist there); nothing under ``oracle/_ref`` is
ever part of the history or of the product.  ``run_train_loop`` then executes the reference's own
``train.train_loop`` with a synthetic ``config`` module injected in ``sys.modules`` (config.py itself cannot be
imported: ``input()`` prompt at config.py:310, hard-coded dataset paths :27-30) - the recipe of SURVEY.md 8(c).

Users: ``oracle/validate_against_reference.py`` (golden vectors), ``bench.py --impl reference`` (CPU arm,
``cpu_baseline.kind = "reference"``), ``__graft_entry__.build()`` (copy recipe only).
"""
from __future__ import annotations

import os
import shutil
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SISR_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("config.py", "model_content_extractor.py", "model_discriminator.py", "model_generator.py",
         "model_generator_progressive.py", "train.py", "utils.py", "visualisation.py")


def make_ref(src: str = REF_SRC, dst: str = REF_DST) -> str | None:
    """Copy the reference's Python files (unmodified) into oracle/_ref.  No-op without ``src``."""
    if not os.path.isdir(src):
        return dst if os.path.exists(os.path.join(dst, "train.py")) else None
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    return dst


def reference_dir() -> str | None:
    """oracle/_ref if it has been made (GPU box, dev container), else /root/reference, else None."""
    for d in (REF_DST, REF_SRC):
        if os.path.exists(os.path.join(d, "train.py")):
            return d
    return None


def import_reference(path: str | None = None):
    """Import the reference's model modules (matplotlib stubbed, vgg19 patched to random init: there is no
    network for the pretrained weights - BASELINE.json prescribes random-init weights)."""
    path = path or reference_dir()
    if path is None:
        raise RuntimeError("reference sources not available (neither oracle/_ref nor /root/reference)")
    if path not in sys.path:
        sys.path.insert(0, path)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torchvision.models as tvm
    if not getattr(tvm, "_sisr_patched", False):
        orig = tvm.vgg19
        tvm.vgg19 = lambda pretrained=True: orig(weights=None)
        tvm._sisr_patched = True
    import model_content_extractor
    import model_discriminator
    import model_generator
    return model_generator, model_discriminator, model_content_extractor


class _Pairs:
    """(hr, dummy) tuples over a list (or any re-iterable) of HR batches, as a DataLoader of an ImageFolder
    yields them (train.py:33)."""

    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        for b in self.batches:
            yield (b, None)


def run_train_loop(net_g, net_d, ext, batches, *, lr, lr_size, device="cpu", content_loss_on_lr=False,
                   loss_weights=None, identity=None, lr_lambda=None, dis_list_old=None, quiet=True):
    """``train.train_loop()`` of the reference, unmodified, on ``batches`` (a list of HR tensors, or of
    ((hr, None), (hr2, None)) pairs in ``content_loss_on_lr`` mode).  The loop stops one batch before the
    end of the list (train.py:35-38), so pass one more batch than steps wanted.  Returns
    (D_losses, G_losses, cont_losses, seconds, optimizers)."""
    import numpy as np
    import torch
    import utils as ref_utils
    ref_utils.save_curr_vis = lambda *a, **k: None
    dev = torch.device(device)
    first = next(iter(getattr(batches, "batches", batches)))
    B = (first[0][0] if content_loss_on_lr else first).shape[0]
    opt_g = torch.optim.Adam(net_g.parameters(), lr=lr, betas=(.9, .999))
    opt_d = torch.optim.Adam(net_d.parameters(), lr=lr, betas=(.9, .999))
    lam = lr_lambda or (lambda it: 1)
    lw = loss_weights or (lambda e: 5e-2, lambda e: 1.0, lambda e: (1.0, ext))
    cfg = types.ModuleType("config")
    cfg.__dict__.update(dict(
        torch=torch, np=np, random=__import__("random"), utils=ref_utils,
        device=dev, net_g=net_g, net_d=net_d, net_content_extractor=ext,
        criterion=torch.nn.BCELoss(), optimizerG=opt_g, optimizerD=opt_d,
        schedulerG=torch.optim.lr_scheduler.LambdaLR(opt_g, lam),
        schedulerD=torch.optim.lr_scheduler.LambdaLR(opt_d, lam),
        real_label=torch.full((B,), 1.0, device=dev), real_label_reduced=torch.full((B,), .9, device=dev),
        fake_label=torch.full((B,), .0, device=dev),
        loss_weight_adv_g=lw[0], loss_weight_adv_d=lw[1], loss_weight_cont=lw[2],
        dataloader_hr=batches if content_loss_on_lr else _Pairs(batches),
        image_size_lr=(3, lr_size, lr_size),
        n_batch=len(batches), num_epochs=1, starting_epoch=0,
        dis_list_old=[] if dis_list_old is None else dis_list_old, dis_list_old_len=1000, dis_list_old_freq=1,
        dis_list_old_ratio=.01, dis_list_old_cpu=True, dis_list_old_save=False,
        content_loss_on_lr=content_loss_on_lr,
        plot_first=False, plot_training=False, plot_usr=False, test_hr=None, test_lr=None,
        write_root="/tmp/", identity=identity))
    sys.modules["config"] = cfg
    sys.modules.pop("train", None)
    import contextlib
    import io
    import train as ref_train
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext():
        d_losses, g_losses, c_losses, _ = ref_train.train_loop()
    return d_losses, g_losses, c_losses, time.perf_counter() - t0, (opt_g, opt_d)
