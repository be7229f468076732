"""Drop-in ``VqAutoEncoder`` (reference: ``vq-video-diffusion/train_vqae.py:22-55``) and its convolutional
encoder / decoder (``autoencoder.py:60-152``), with the reference's attribute names so that the authors'
checkpoints load with ``load_state_dict(strict=True)``.

The convolution stacks are stock PyTorch / cuDNN (out of scope for hand-written kernels); what
this module adds to the hot path is the glue around the quantizer: frames -> latents (BCHW -> BHWC) ->
``wm_vq_nearest`` -> token grid, and tokens -> codebook gather -> BCHW -> decoder.  The denoiser trainer
(``main.py:234-237``) and the sampler (``main.py:113``) call exactly these two functions.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .vq import VectorQuantizerEMA


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2 if k == 3 else 0, bias=bias)


class Residual(nn.Module):
    """conv3x3(stride) - BN - LeakyReLU - conv1x1 - BN, plus a skip that is strided by a ``stride x stride``
    convolution + BN when ``stride != 1`` (``autoencoder.py:18-42``).  Sequential indices 0,1,3,4 carry the
    parameters (index 2 is the activation)."""

    def __init__(self, planes, hidden_planes, stride=1):
        super().__init__()
        self._block = nn.Sequential(_conv(planes, hidden_planes, 3, stride), nn.BatchNorm2d(hidden_planes),
                                    nn.LeakyReLU(inplace=True), _conv(hidden_planes, planes, 1), nn.BatchNorm2d(planes))
        self.downsample = None
        if stride != 1:
            self.downsample = nn.Sequential(nn.Conv2d(planes, planes, kernel_size=stride, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes))

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        return F.leaky_relu(self._block(x) + skip)


class ResidualStack(nn.Module):
    """``num_layers`` x (Residual stride 1, Residual stride 2): every layer halves H and W (``:45-57``)."""

    def __init__(self, in_planes, num_layers, hidden_planes):
        super().__init__()
        self._num_residual_layers = num_layers
        self._stack = nn.Sequential(*[Residual(in_planes, hidden_planes, stride)
                                      for _ in range(num_layers) for stride in (1, 2)])

    def forward(self, x):
        return self._stack(x)


class SimpleResidualEncoder(nn.Module):
    def __init__(self, in_planes, out_planes, num_layers, hidden_planes):
        super().__init__()
        self._conv_1 = _conv(in_planes, out_planes, 3)
        self._residual_stack = ResidualStack(out_planes, num_layers, hidden_planes)

    def forward(self, x):
        return self._residual_stack(F.leaky_relu(self._conv_1(x)))


class UpscaleResidual(nn.Module):
    """BN - LeakyReLU - (x2 bilinear) - conv3x3 - BN - LeakyReLU - conv3x3, plus a 1x1-conv skip over the
    (upsampled) input (``autoencoder.py:89-131``)."""

    def __init__(self, in_planes, out_planes, upsample=True):
        super().__init__()
        self.conv1 = _conv(in_planes, out_planes, 3, bias=True)
        self.conv2 = _conv(out_planes, out_planes, 3, bias=True)
        self.bn1 = nn.BatchNorm2d(in_planes)
        self.bn2 = nn.BatchNorm2d(out_planes)
        self.upsample = upsample
        self.learn_conv_residual = in_planes != out_planes or upsample
        if self.learn_conv_residual:
            self.conv_residual = nn.Conv2d(in_planes, out_planes, kernel_size=1)

    @staticmethod
    def _up(t):
        return F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False)

    def forward(self, x):
        h = F.leaky_relu(self.bn1(x))
        if self.upsample:
            h, x = self._up(h), self._up(x)
        h = self.conv2(F.leaky_relu(self.bn2(self.conv1(h))))
        return h + (self.conv_residual(x) if self.learn_conv_residual else x)


class SimpleResidualDecoder(nn.Module):
    def __init__(self, cfg, in_channels, out_channels=3):
        super().__init__()
        stages = [_conv(in_channels, in_channels, 3)]
        for width in cfg:
            stages.append(UpscaleResidual(in_channels, width, True))
            in_channels = width
        stages.append(_conv(in_channels, out_channels, 3))
        self.decoder_stack = nn.Sequential(*stages)

    def forward(self, x):
        return self.decoder_stack(x)


class VqAutoEncoder(nn.Module):
    """``encoder -> VectorQuantizerEMA -> decoder`` with the reference's constructor and three entry points."""

    def __init__(self, embedding_dim, num_embeddings, downscale_steps=2, hidden_planes=128, in_channels=3):
        super().__init__()
        self.encoder = SimpleResidualEncoder(in_channels, embedding_dim, downscale_steps, hidden_planes)
        self.decoder = SimpleResidualDecoder([hidden_planes] * downscale_steps, in_channels=embedding_dim,
                                             out_channels=in_channels)
        self.vq = VectorQuantizerEMA(embedding_dim, num_embeddings)

    def forward(self, x):
        """``-> (reconstruction, latent_loss, perplexity)`` (``train_vqae.py:33-43``)."""
        h = self.encoder(x).permute(0, 2, 3, 1)                         # BCHW -> BHWC: the quantizer's layout
        quantized, _, latent_loss, perplexity = self.vq(h)
        return self.decoder(quantized.permute(0, 3, 1, 2).contiguous()), latent_loss, perplexity

    @torch.no_grad()
    def encode(self, x):
        """Frames ``[B,C,H,W]`` -> token grid ``int64 [B,h,w]`` (``:45-49``)."""
        h = self.encoder(x).permute(0, 2, 3, 1)
        return self.vq.encode(h).view(h.shape[:-1])

    @torch.no_grad()
    def decode(self, z):
        """Token grid ``[B,h,w]`` -> frames ``[B,C,H,W]`` (``:51-55``)."""
        return self.decoder(self.vq.decode(z).permute(0, 3, 1, 2))
