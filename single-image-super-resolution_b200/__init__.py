"""sisr_b200 - B200-native SRGAN training step (hot path of keyber/Single-Image-Super-Resolution).

Drop-in modules: ``model_generator.Generator`` / ``GeneratorSuffix``,
``model_discriminator.Discriminator``, ``model_content_extractor.MaskedVGG`` - same constructors,
attributes and ``state_dict`` keys as the reference - running on hand-written sm_100a CUDA kernels
behind the C ABI of ``include/sisr_b200.h``.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .model_content_extractor import MaskedVGG, identity
from .model_discriminator import Discriminator
from .model_generator import Generator, GeneratorSuffix
from .optim import Adam
from .train import SRGANTrainer, StepConfig
from .utils import lr_from_hr
from .metrics import evaluate, psnr_ssim

__all__ = ["Generator", "GeneratorSuffix", "Discriminator", "MaskedVGG", "identity", "Adam",
           "SRGANTrainer", "StepConfig", "lr_from_hr", "psnr_ssim", "evaluate"]
