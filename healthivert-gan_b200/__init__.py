"""healthivert-gan_b200: B200-native two-stage pseudo-healthy synthesis path of HealthiVert-GAN.

Import as ``healthivert_gan_b200`` (the hyphenated directory is not a Python identifier; the
sibling ``healthivert_gan_b200/`` package forwards here).
"""
from . import _lib  # noqa: F401
from .inpaint_networks import (Conv2dBlock, ContextualAttention, CoarseGenerator, FineGenerator,  # noqa: F401
                               Generator, gen_conv)
from .edge_operator import Sobel, edge_mse_loss  # noqa: F401
from . import mask_ops  # noqa: F401
from . import sharding  # noqa: F401
from .pipeline import SlicePipeline  # noqa: F401

__all__ = ["Generator", "CoarseGenerator", "FineGenerator", "ContextualAttention", "Conv2dBlock", "gen_conv",
           "Sobel", "edge_mse_loss", "mask_ops", "sharding", "SlicePipeline"]
