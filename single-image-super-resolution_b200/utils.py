"""Device-side replacement of the one ``utils`` function on the training path (reference:
utils.py:16-31); plotting / checkpoint helpers of the reference's ``utils`` are out of scope.

    import utils, sisr_b200.utils
    utils.lr_from_hr = sisr_b200.utils.lr_from_hr      # train.py:46 then runs the CUDA kernel
"""
from __future__ import annotations

from . import ops


def lr_from_hr(img_hr, image_size_lr, device="cpu"):
    """hr in [-1, 1] -> clamp(bicubic(hr)) of size ``image_size_lr`` (same signature as the
    reference; ``device`` is accepted and ignored: the result lives where ``img_hr`` lives)."""
    if isinstance(image_size_lr, int):
        image_size_lr = (image_size_lr, image_size_lr)
    return ops.LrFromHrFn.apply(img_hr, tuple(image_size_lr))
