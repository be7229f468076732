"""B200-native local-3D-attention / VQ denoiser hot path (drop-in for world-modelz).

Public surface mirrors the reference modules this package replaces:

* ``Local3dAttention``, ``Local3dAttentionTransformer``, ``PreNorm``, ``FeedForward``
  (``vq-video-diffusion/local_3d_attention.py``)
* ``VectorQuantizerEMA`` (``vq-video-diffusion/vq.py``)
* ``VqVideoDiffusionModel`` (``vq-video-diffusion/main.py:25-36``)
* ``VqAutoEncoder`` (``vq-video-diffusion/train_vqae.py:22-55``)

plus the data-parallel trainer / sampler built on them (``DenoiserTrainer``,
``sample_next_frame``, ``sample_frames``, ``LossAwareSamplerEma``).  All hot-path arithmetic runs in ``libwm_b200.so``
(``include/wm_b200.h``); importing the ops without the built library raises.
"""
from .local_3d_attention import FeedForward, Local3dAttention, Local3dAttentionTransformer, PreNorm
from .vq import VectorQuantizerEMA
from .vqae import VqAutoEncoder
from .denoiser import (DenoiserTrainer, LossAwareSamplerEma, VqVideoDiffusionModel, corrupt_last_frame, sample_frames,
                       sample_next_frame)

__all__ = ['PreNorm', 'FeedForward', 'Local3dAttention', 'Local3dAttentionTransformer', 'VectorQuantizerEMA',
           'VqAutoEncoder', 'VqVideoDiffusionModel', 'DenoiserTrainer', 'LossAwareSamplerEma', 'corrupt_last_frame',
           'sample_next_frame', 'sample_frames']
