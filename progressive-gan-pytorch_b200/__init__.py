"""progan_b200 — B200-native progressive-GAN training step (G/D forward+backward + WGAN-GP).

Drop-in for the hot path of gwilczynski95/Progressive-GAN-pytorch:
    from progan_b200 import Generator, Discriminator      # instead of progan_modules
"""
from .kernels import ConvOp, get_kernels, set_kernels
from .progan_modules import (ConditionalCorrectDiscriminatorAda, ConditionalCorrectDiscriminatorWgangp,
                             ConditionalCorrectGenerator, ConditionalCorrectGeneratorAda,
                             ConditionalDiscriminatorWgangp, ConditionalGenerator, ConvBlock,
                             CorrectDiscriminator, CorrectGenerator, Discriminator, EqualEmbed,
                             EqualConv2d, EqualConvTranspose2d, EqualLinear, Generator, MnistConvBlock,
                             PixelNorm,
                             set_default_precision)
from .functions import gradient_penalty
from .train import MiniStepSchedule, ProgressiveSchedule, Trainer
from .ada import AdaptiveAugment, AugmentPipe
from . import ada, mnist_pggan

__all__ = ["Generator", "Discriminator", "CorrectGenerator", "CorrectDiscriminator", "mnist_pggan", "ConditionalCorrectGenerator",
           "ConditionalCorrectDiscriminatorWgangp", "ConditionalGenerator", "ConditionalDiscriminatorWgangp",
           "ConditionalCorrectGeneratorAda", "ConditionalCorrectDiscriminatorAda", "EqualEmbed", "ConvBlock", "EqualConv2d", "EqualConvTranspose2d",
           "EqualLinear", "PixelNorm", "ConvOp", "get_kernels", "set_kernels",
           "set_default_precision", "gradient_penalty", "Trainer", "ProgressiveSchedule", "MiniStepSchedule",
           "AugmentPipe", "AdaptiveAugment", "ada"]
