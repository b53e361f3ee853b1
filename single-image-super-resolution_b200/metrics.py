"""PSNR / SSIM on the device (the reference's README.md:88 todo) and the viewer flow of
``visualisation.py`` (visualisation.py:16-52) without the plotting: LR = lr_from_hr(HR), SR = G(LR),
UR = G(HR) in eval mode, plus the quality of SR against HR."""
from __future__ import annotations

import torch

from . import ops
from ._lib import call
from .utils import lr_from_hr


def psnr_ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0):
    """Per-image PSNR [dB] and SSIM (11x11 Gaussian window, sigma 1.5, K1 = 0.01, K2 = 0.03, valid
    positions, mean over channels) of two NCHW fp32 batches; ``data_range`` = 2 for images in [-1, 1]."""
    ops._require_cuda(a, "psnr_ssim")
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError(f"psnr_ssim: shapes {tuple(a.shape)} and {tuple(b.shape)} must be equal NCHW")
    a, b = a.contiguous().float(), b.contiguous().float()
    n, c, h, w = a.shape
    ws = torch.empty(2 * n, dtype=torch.float32, device=a.device)
    psnr = torch.empty(n, dtype=torch.float32, device=a.device)
    ssim = torch.empty(n, dtype=torch.float32, device=a.device)
    call("sisr_psnr_ssim", a, b, n, c, h, w, float(data_range), ws, psnr, ssim, ops._stream())
    return psnr, ssim


@torch.no_grad()
def evaluate(net_g, hr: torch.Tensor, image_size_lr):
    """visualisation.py:46-52: ``lr = utils.lr_from_hr(hr, size); sr = net_g(lr); ur = net_g(hr)`` with the
    generator in eval mode (BN running statistics, no spectral-norm iteration), and PSNR / SSIM of SR
    against HR.  Returns {"lr", "sr", "hr", "ur", "psnr", "ssim"}."""
    was_training = net_g.training
    net_g.eval()
    try:
        lr = lr_from_hr(hr, image_size_lr)
        sr = net_g(lr)
        ur = net_g(hr)
    finally:
        net_g.train(was_training)
    out = {"lr": lr, "sr": sr, "hr": hr, "ur": ur}
    if sr.shape == hr.shape and min(hr.shape[-2:]) >= 11:
        out["psnr"], out["ssim"] = psnr_ssim(sr.clamp(-1, 1), hr)
    return out
