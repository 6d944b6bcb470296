"""B200-native generate hot path of GAN-segmentation: StyleGAN-v1 synthesis + segmentation
decoder -> image + argmax mask.  Host mirror of the reference's Python surface
(ImageGenerator / SegSolver / Generator / Decoder) over a C-ABI CUDA library (csrc/, include/gsx.h).
"""
from .config import generator_config, decoder_config, MAX_RES_LOG2  # noqa: F401

__all__ = ['generator_config', 'decoder_config', 'MAX_RES_LOG2']
