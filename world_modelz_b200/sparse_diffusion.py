"""Sparse-context denoiser of ``minecraft/sparse_diffusion.py``: the model sees ``num_context`` (512) token positions
gathered from the (S,H,W) grid instead of the whole clip, runs a dense transformer over them and predicts every
gathered position.  Drop-in for ``VqSparseDiffusionModel`` (``sparse_diffusion.py:75-111``) and the ViT-style encoder it
wraps (``minecraft/transformer.py:34-80``), with the reference's ``state_dict`` keys so that the authors' ``.pth`` files
load with ``strict=True``.

The 512-token attention is dense, so it runs on stock scaled-dot-product attention (out of scope for
a hand-written kernel); what this module takes from the hot path is the block schedule around it: residual add +
LayerNorm with the deferred biases, bias + GELU (``wm_add_layernorm_*``, ``wm_bias_gelu_*``), and the position samplers
on the device without per-sample Python loops.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .local_3d_attention import FeedForward, PreNorm


def sample_flat_positions(batch_size, context_length, s, h, w, device):
    """``[batch_size, context_length]`` positions: consecutive random permutations of the grid, cut into rows
    (``sparse_diffusion.py:31-41``) -- positions are unique inside every permutation-sized run."""
    max_index = s * h * w
    n = batch_size * context_length
    runs = (n + max_index - 1) // max_index
    keys = torch.rand(runs, max_index, device=device)
    return keys.argsort(dim=1).reshape(-1)[:n].view(batch_size, context_length)


def sample_time_dependent(batch_size, context_length, s, h, w, t, device, o=None):
    """Positions drawn uniformly WITHOUT replacement from a window of whole frames whose length grows with the
    diffusion time ``t`` and whose start is ``o`` (random when None) -- ``sparse_diffusion.py:44-72``, with the
    per-sample ``randperm`` loop replaced by one batched random-key selection on the device."""
    t = t.reshape(-1).clamp(0, 1).to(device)
    if context_length <= 0:
        raise ValueError('context_length must be positive')
    min_sample_window = math.ceil(context_length / (h * w))
    if not min_sample_window < s:
        raise ValueError('context does not fit into fewer frames than the clip has')
    window = torch.floor(min_sample_window + (t * (s - min_sample_window + 1))).clamp(max=s - min_sample_window)
    o = torch.rand_like(t) if o is None else o.reshape(-1).clamp(0, 1 - 1e-5).to(device)
    offset = torch.floor(o * (s - window + 1)).long() * (h * w)
    count = window.long() * (h * w)                                           # positions in each sample's window
    keys = torch.rand(batch_size, s * h * w, device=device)
    keys = keys.masked_fill(torch.arange(s * h * w, device=device)[None, :] >= count[:, None], 2.0)
    return keys.topk(context_length, dim=1, largest=False).indices + offset[:, None]


class Attention(nn.Module):
    """Dense multi-head attention with a fused ``to_qkv`` projection (``minecraft/transformer.py:34-64``)."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        if heads == 1 and dim_head == dim:
            self.to_out = nn.Identity()
        else:
            self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def _core(self, x):
        B, n, _ = x.shape
        q, k, v = (t.view(B, n, self.heads, -1).transpose(1, 2) for t in self.to_qkv(x).chunk(3, dim=-1))
        out = F.scaled_dot_product_attention(q, k, v, dropout_p=self.dropout.p if self.training else 0.0, scale=self.scale)
        return out.transpose(1, 2).reshape(B, n, -1)

    def forward(self, x):
        return self.to_out(self._core(x))

    def forward_deferred_bias(self, x):
        """``(y, bias)`` with ``forward(x) == y + bias``: ``to_out.0``'s bias is applied by the caller's fused
        residual-add + LayerNorm kernel (whose backward then reduces its gradient); ``bias`` None when it cannot be."""
        core = self._core(x)
        if isinstance(self.to_out, nn.Identity) or (self.training and self.to_out[1].p > 0.0):
            return self.to_out(core), None
        return F.linear(core, self.to_out[0].weight), self.to_out[0].bias


class Transformer(nn.Module):
    """``depth`` x (PreNorm attention, PreNorm MLP) residual blocks (``minecraft/transformer.py:67-80``), every residual
    add fused with the LayerNorm that follows it."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList(
            nn.ModuleList([PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                           PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout))]) for _ in range(depth))

    def forward(self, x):
        pending = pending_bias = None
        for attn, ff in self.layers:
            x, xn = ops.add_layernorm(x, pending, attn.norm.weight, attn.norm.bias, attn.norm.eps, pending_bias)
            a, a_bias = attn.fn.forward_deferred_bias(xn)
            x, xn = ops.add_layernorm(x, a, ff.norm.weight, ff.norm.bias, ff.norm.eps, a_bias)
            pending, pending_bias = ff.fn.forward_deferred_bias(xn)
        if pending is None:
            return x
        return x + pending if pending_bias is None else x + (pending + pending_bias.to(pending.dtype))


class VqSparseDiffusionModel(nn.Module):
    """``forward(x [B,n] tokens, indices [B,n] flat grid positions) -> logits [B,n,num_classes]``."""

    def __init__(self, *, shape, dim, num_classes, depth, dim_head, mlp_dim, heads=1, dropout=0.0):
        super().__init__()
        self.shape = shape
        S, H, W = shape
        self.pos_emb_s = nn.Embedding(S, dim)
        self.pos_emb_h = nn.Embedding(H, dim)
        self.pos_emb_w = nn.Embedding(W, dim)
        self.embedding = nn.Embedding(num_classes + 1, dim)                   # +1: the mask token
        self.transformer = Transformer(dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim,
                                       dropout=dropout)
        self.logit_proj = nn.Linear(dim, num_classes)

    def pos_embedding_3d(self, indices):
        S, H, W = self.shape
        w_pos = indices % W
        h_pos = indices.div(W, rounding_mode='trunc') % H
        s_pos = indices.div(H * W, rounding_mode='trunc')
        return self.pos_emb_s(s_pos) + self.pos_emb_h(h_pos) + self.pos_emb_w(w_pos)

    def forward(self, x, indices):
        h = self.embedding(x) + self.pos_embedding_3d(indices)
        return self.logit_proj(self.transformer(h))
