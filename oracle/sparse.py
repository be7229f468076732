"""Oracle (CPU restatement) of the sparse-context denoiser, ``minecraft/sparse_diffusion.py:75-111`` over the dense
transformer of ``minecraft/transformer.py:34-80``.  TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Functional over a flat dict with the reference's ``state_dict`` keys; plain fp32 PyTorch on the CPU, written as the
textbook formulas (explicit softmax(QK^T)V, LayerNorm, GELU) rather than as the modules.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def sparse_denoiser_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, indices: torch.Tensor, shape, depth: int,
                            heads: int) -> torch.Tensor:
    """``x, indices [B,n]`` -> logits ``[B,n,K]`` (``sparse_diffusion.py:107-111``)."""
    S, H, W = shape
    w_pos = indices % W
    h_pos = indices.div(W, rounding_mode='trunc') % H
    s_pos = indices.div(H * W, rounding_mode='trunc')
    h = p['embedding.weight'][x] + p['pos_emb_s.weight'][s_pos] + p['pos_emb_h.weight'][h_pos] + p['pos_emb_w.weight'][w_pos]
    dim = h.shape[-1]
    B, n, _ = h.shape
    for i in range(depth):
        pre = f'transformer.layers.{i}.'
        xn = F.layer_norm(h, (dim,), p[pre + '0.norm.weight'], p[pre + '0.norm.bias'])
        qkv = xn @ p[pre + '0.fn.to_qkv.weight'].t()
        q, k, v = (t.view(B, n, heads, -1).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        d = q.shape[-1]
        attn = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)          # transformer.py:56-58
        out = (attn @ v).transpose(1, 2).reshape(B, n, heads * d)
        if pre + '0.fn.to_out.0.weight' in p:
            out = out @ p[pre + '0.fn.to_out.0.weight'].t() + p[pre + '0.fn.to_out.0.bias']
        h = out + h
        xn = F.layer_norm(h, (dim,), p[pre + '1.norm.weight'], p[pre + '1.norm.bias'])
        m = F.gelu(xn @ p[pre + '1.fn.net.0.weight'].t() + p[pre + '1.fn.net.0.bias'])
        h = m @ p[pre + '1.fn.net.3.weight'].t() + p[pre + '1.fn.net.3.bias'] + h
    return h @ p['logit_proj.weight'].t() + p['logit_proj.bias']
