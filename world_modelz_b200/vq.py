"""Drop-in ``VectorQuantizerEMA`` (reference: ``vq-video-diffusion/vq.py:6-111``).

Same constructor, buffers (``embedding``/``cluster_size`` persistent; ``latent_offsets``,
``activation_count``, ``accumulated_error`` non-persistent), flags and 4-tuple return.
The distance / argmin / gather / per-latent error step -- the part that materialises an
``[N,L,D,K]`` temporary in the reference -- runs in ``wm_vq_nearest``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _StraightThrough(torch.autograd.Function):
    """forward: the kernel's ``x + (e - x)`` value; backward: identity to ``x`` (``vq.py:70``)."""

    @staticmethod
    def forward(ctx, x, ste_value):
        return ste_value.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g, None


class VectorQuantizerEMA(nn.Module):
    def __init__(self, embedding_dim, num_embeddings, num_latents=1, decay=0.99, eps=1e-5):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.num_latents = num_latents
        self.decay = decay
        self.eps = eps
        L, K, D = num_latents, num_embeddings, embedding_dim
        self.register_buffer('embedding', torch.randn(L, K, D))
        self.register_buffer('cluster_size', torch.ones(L, K))
        self.register_buffer('latent_offsets', (torch.arange(L) * K).unsqueeze(0), persistent=False)
        self.register_buffer('activation_count', torch.zeros(L, K), persistent=False)
        self.register_buffer('accumulated_error', torch.zeros(L, K), persistent=False)
        self.simple_update = False
        self.laplace_smoothing = True

    # -- pieces ------------------------------------------------------------------------
    def _flat(self, input):
        return input.reshape(-1, self.num_latents, self.embedding_dim)

    def codebook_distance(self, input, normalize=True):
        return ops.vq_distance(self._flat(input), self.embedding, normalize)

    def encode(self, input):
        idx, _, _ = ops.vq_nearest(self._flat(input).float(), self.embedding, want_quantized=False, want_err=False)
        return idx

    def decode(self, indices):
        flat = (self.latent_offsets + indices.reshape(-1, self.num_latents)).reshape(-1)
        codes = self.embedding.reshape(self.num_latents * self.num_embeddings, self.embedding_dim)
        return codes.index_select(0, flat).reshape(*indices.shape, self.embedding_dim)

    # -- forward -----------------------------------------------------------------------
    def forward(self, input):
        """``-> (quantized [straight-through], encodings one-hot f32 [N,L,K], commitment_loss, perplexity)``
        (``vq.py:25-75``).  Nearest code, gather and per-latent error come from ``wm_vq_nearest``; the per-code
        statistics (``accumulated_error`` always, ``activation_count`` / ``cluster_size`` / ``embedding`` EMA in
        training mode) from ONE ``wm_vq_stats`` pass over the indices -- no ``[N,L,D,K]`` distance temporary and no
        one-hot matmul; the one-hot the API returns is written directly by ``wm_vq_onehot``."""
        L, K, D = self.num_latents, self.num_embeddings, self.embedding_dim
        flat = self._flat(input)
        x32 = flat.detach().float().contiguous()
        idx, ste, err = ops.vq_nearest(x32, self.embedding)
        n = flat.shape[0]
        with torch.no_grad():
            counts = torch.zeros(L, K, device=flat.device, dtype=torch.float32)
            dw = torch.zeros(L, K, D, device=flat.device, dtype=torch.float32) if self.training else None
            ops.vq_stats(x32, idx, err, counts, dw, self.accumulated_error)          # vq.py:35-36, 43-46
            encodings = ops.vq_onehot(idx, K)                                        # vq.py:39
            quantized = self.decode(idx).reshape(input.shape)
            if self.training:
                self._ema_update(counts, dw)
        commitment_loss = F.mse_loss(quantized, input)
        out = _StraightThrough.apply(input, ste.to(input.dtype))
        avg = counts / n
        perplexity = torch.exp(-torch.sum(avg * torch.log(avg + 1e-10) / L))
        return out, encodings, commitment_loss, perplexity

    def _ema_update(self, counts, dw):
        """Cluster-size / embedding EMA with Laplace smoothing (``vq.py:47-65``) from the fused statistics."""
        K = self.num_embeddings
        self.activation_count.add_(counts)
        if self.simple_update:
            dw = dw / counts.unsqueeze(-1)
            ok = dw == dw
            self.embedding.data[ok] = self.decay * self.embedding.data[ok] + (1.0 - self.decay) * dw[ok]
            return
        self.cluster_size.mul_(self.decay).add_(counts, alpha=1 - self.decay)
        if self.laplace_smoothing:
            tot = self.cluster_size.sum(dim=-1, keepdim=True)
            size = (self.cluster_size + self.eps) / (tot + K * self.eps) * tot
        else:
            size = self.cluster_size
        self.embedding.mul_(self.decay).add_(dw / size.unsqueeze(-1), alpha=1 - self.decay)

    # -- housekeeping --------------------------------------------------------------------
    def reuse_inactive(self):
        total = 0
        for l in range(self.num_latents):
            dead = self.activation_count[l] == 0
            nd = int(dead.count_nonzero())
            if nd > 0:
                _, top = self.activation_count[l].topk(nd)
                self.embedding[l][dead] = self.embedding[l][dead] * 0.1 + self.embedding[l][top] * 0.9
                total += nd
        return total

    def reset_stats(self):
        self.activation_count.zero_()
        self.accumulated_error.zero_()
