"""Drop-in ``local_3d_attention`` module: same classes, constructor arguments, forward
signatures and ``state_dict`` keys as ``vq-video-diffusion/local_3d_attention.py``; the
attention core runs in the fused sm_100a kernels behind ``ops.local3d_attention``.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class PreNorm(nn.Module):
    """LayerNorm on the positional input only; keyword inputs (``q=``) bypass the norm
    (reference ``local_3d_attention.py:11-17``)."""

    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)

    def forward_prenormed(self, x_normed, **kwargs):
        """Run the wrapped module on an input the caller has already normalised with ``self.norm``'s
        parameters (the transformer fuses that LayerNorm with the preceding residual add)."""
        return self.fn(x_normed, **kwargs)


class FeedForward(nn.Module):
    """Linear-GELU-Linear with the reference's Sequential indices (``net.0`` / ``net.3``)."""

    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        stages = [nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                  nn.Linear(hidden_dim, dim), nn.Dropout(dropout)]
        self.net = nn.Sequential(*stages)

    def _no_dropout(self):
        return not self.training or (self.net[2].p == 0.0 and self.net[4].p == 0.0)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('world_modelz_b200 modules run on a CUDA device only (no CPU fallback)')
        if self._no_dropout():
            h = ops.bias_gelu(torch.nn.functional.linear(x, self.net[0].weight), self.net[0].bias)
            return ops.linear(h, self.net[3].weight, self.net[3].bias)
        return self.net(x)                # dropout active: stock modules, still on the device

    def forward_deferred_bias(self, x):
        """``(y, bias)`` with ``forward(x) == y + bias``: the caller adds ``net.3``'s bias inside its fused
        residual-add + LayerNorm kernel, whose backward then also reduces the bias gradient.  ``bias`` is None when the
        bias could not be deferred (dropout active)."""
        if x.is_cuda and self._no_dropout():
            h = ops.bias_gelu(torch.nn.functional.linear(x, self.net[0].weight), self.net[0].bias)
            return torch.nn.functional.linear(h, self.net[3].weight), self.net[3].bias
        return self.forward(x), None


class Local3dAttention(nn.Module):
    """NUWA-style "nearby" attention over a (S,H,W) token grid.

    ``forward(x, q)``: K and V are projected from ``x``, Q from ``q``; each query attends to
    the ``(2e+1)^3`` window around it, neighbours outside the grid excluded.  Parameters:
    ``to_q/to_k`` (no bias), ``to_v`` (bias), ``to_out.0`` unless ``heads == 1 and
    dim_head == dim`` (reference ``:35-55``).  fp32 tensors take the exact SIMT kernels,
    bf16 tensors the tcgen05 kernels.  ``use_checkpointing`` is accepted for API
    compatibility; the fused kernels keep only O and the LSE, so there is nothing to
    recompute.
    """

    def __init__(self, extents, dim, heads=8, dim_head=64, dropout=.0, use_checkpointing=True):
        super().__init__()
        if len(extents) != 3:
            raise ValueError('extents must be (S, H, W)')
        self.extents = extents
        self.heads = heads
        self.scale = dim_head ** -0.5
        inner_dim = heads * dim_head
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_k = nn.Linear(dim, inner_dim, bias=False)
        self.to_v = nn.Linear(dim, inner_dim, bias=True)
        if heads == 1 and dim_head == dim:
            self.to_out = nn.Identity()
        else:
            self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))
        self.use_checkpointing = use_checkpointing
        self.kernel_flags = 0          # ops.FLAG_SIMT forces the SIMT kernels for bf16 (device cross-check)

    def local_attention(self, k, v, q):
        """Reference-shaped core: ``[B,S,H,W,heads*d]`` in, ``[(B S H W), heads, 1, d]`` out."""
        out = ops.local3d_attention(q, k, v, self.heads, self.extents, self.scale, self.kernel_flags)
        return out.reshape(-1, self.heads, 1, out.shape[-1] // self.heads)

    def forward_deferred_bias(self, x, q, q_projected=None, drop_frames=0):
        """``(y, bias)`` with ``forward(x, q) == y + bias`` (``q_projected``: ``to_q(q)`` when the caller already has it;
        ``drop_frames``: return ``y[:, drop_frames:]`` only -- the output projection then runs on those frames alone).  Softmax rows sum to one, so ``to_v``'s bias passes through
        the attention core unchanged: ``attn(q, k, v + b_v) = attn(q, k, v) + b_v``, and with the output projection
        the whole module equals ``attn(q, k, x W_v^T) W_o^T + (W_o b_v + b_o)``.  The caller adds that [dim] vector in
        its fused residual-add + LayerNorm kernel; autograd routes its gradient (reduced by that kernel's backward) to
        ``to_v.bias``, ``to_out.0.bias`` and ``to_out.0.weight``.  ``bias`` is None when this does not apply
        (no output projection, dropout active, CPU tensors)."""
        deferrable = (x.is_cuda and isinstance(self.to_out, nn.Sequential)
                      and (not self.training or self.to_out[1].p == 0.0))
        if not deferrable:
            return self.forward(x, q)[:, drop_frames:], None
        if x.dim() != 5 or q.shape[:-1] != x.shape[:-1]:
            raise ValueError(f'expected x, q of shape [B,S,H,W,dim], got {tuple(x.shape)} and {tuple(q.shape)}')
        w_o = self.to_out[0].weight
        # to_k and to_v read the same normalised input: ONE GEMM with N = 2*inner; the kernels take the two channel
        # halves of its output in place, and backward hands dK | dV back as one operand (one dgrad, one wgrad)
        kv = torch.nn.functional.linear(x, torch.cat((self.to_k.weight, self.to_v.weight), dim=0))
        qp = q_projected if q_projected is not None else self.to_q(q)
        core = ops.local3d_attention_kv(qp, kv, self.heads, self.extents, self.scale, self.kernel_flags)
        bias = torch.addmv(self.to_out[0].bias, w_o, self.to_v.bias)      # W_o b_v + b_o, one GEMV
        if drop_frames:
            core = core.reshape(*q.shape[:-1], -1)[:, drop_frames:]
            return torch.nn.functional.linear(core, w_o), bias
        return torch.nn.functional.linear(core, w_o).reshape(q.shape), bias

    def forward(self, x, q):
        if x.dim() != 5 or q.shape[:-1] != x.shape[:-1]:
            raise ValueError(f'expected x, q of shape [B,S,H,W,dim], got {tuple(x.shape)} and {tuple(q.shape)}')
        v = ops.linear(x, self.to_v.weight, self.to_v.bias)
        core = ops.local3d_attention(self.to_q(q), self.to_k(x), v, self.heads, self.extents, self.scale,
                                     self.kernel_flags)
        if isinstance(self.to_out, nn.Identity):
            return core.reshape(q.shape)
        out = ops.linear(core, self.to_out[0].weight, self.to_out[0].bias)
        return self.to_out[1](out).reshape(q.shape)


class Local3dAttentionTransformer(nn.Module):
    """Token embedding + three axis position embeddings + ``depth`` x (attention, MLP)
    residual blocks (reference ``local_3d_attention.py:121-163``)."""

    def __init__(self, *, data_shape, dim, num_classes, extents, depth, heads, dim_head, mlp_dim, dropout=.0):
        super().__init__()
        self.num_classes = num_classes
        self.embedding = nn.Embedding(num_classes, dim)
        self.pos_emb_s = nn.Embedding(data_shape[0], dim)
        self.pos_emb_h = nn.Embedding(data_shape[1], dim)
        self.pos_emb_w = nn.Embedding(data_shape[2], dim)
        self.layers = nn.ModuleList(
            nn.ModuleList([
                PreNorm(dim, Local3dAttention(extents, dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
            ]) for _ in range(depth))

    def get_pos_embedding(self, batch_shape):
        _, s, h, w = batch_shape
        pos = (self.pos_emb_s.weight[:s, None, None, :] + self.pos_emb_h.weight[None, :h, None, :]
               + self.pos_emb_w.weight[None, None, :w, :])
        return pos.unsqueeze(0).expand(batch_shape[0], -1, -1, -1, -1)

    def last_frame_cone(self, s):
        """``lo[i]``: the first frame the input of layer ``i`` has to cover for the LAST frame of the transformer's
        output to be exact (``lo[depth] = s - 1``).  A query sees ``extents[0]`` frames on either side, so the last
        frame depends on ``depth * extents[0]`` frames before it and on nothing earlier: frames below ``lo[i]`` are
        dead inputs of layer ``i`` for ``forward(img_z)[:, -1]`` -- their contribution to it and to its gradient is
        exactly zero."""
        depth = len(self.layers)
        e_s = self.layers[0][0].fn.extents[0] if depth else 0
        return [max(0, s - 1 - (depth - i) * e_s) for i in range(depth + 1)]

    def _blocks(self, img_z, cone=None):
        """The residual stream before the last MLP branch is added, that branch's output and its deferred bias.  With a
        ``cone`` (``last_frame_cone``) only frames ``>= cone[i]`` enter layer ``i``: the attention of a slice that does
        not start at frame 0 is wrong on its first ``extents[0]`` frames (their earlier neighbours are missing), which
        are exactly the frames ``cone[i + 1]`` drops before the branch is added."""
        _, s, h, w = img_z.shape
        lo = cone[0] if cone is not None else 0
        if lo:
            img_z = img_z[:, lo:].contiguous()
        x = ops.embed_pos(img_z, self.embedding.weight, self.pos_emb_s.weight[lo:s], self.pos_emb_h.weight[:h],
                          self.pos_emb_w.weight[:w])
        pending = pending_bias = None                # branch output (and its deferred bias) not yet added to the stream
        for i, (attn, ff) in enumerate(self.layers):
            x, xn = ops.add_layernorm(x, pending, attn.norm.weight, attn.norm.bias, attn.norm.eps, pending_bias)
            # to_q reads the residual stream itself (reference :160, q=x): the stream goes THROUGH the projection node, so
            # that its gradient is folded into the projection's dgrad GEMM instead of a separate full-size add
            qp, x = ops.linear_passthrough(x, attn.fn.to_q.weight)
            drop = cone[i + 1] - cone[i] if cone is not None else 0
            a, a_bias = attn.fn.forward_deferred_bias(xn, q=x, q_projected=qp, drop_frames=drop)
            if drop:
                x, a = x[:, drop:].contiguous(), a.contiguous()
            x, xn = ops.add_layernorm(x, a, ff.norm.weight, ff.norm.bias, ff.norm.eps, a_bias)
            pending, pending_bias = ff.fn.forward_deferred_bias(xn)
        return x, pending, pending_bias

    @staticmethod
    def _close(x, pending, pending_bias):
        if pending is None:
            return x
        return x + pending if pending_bias is None else x + (pending + pending_bias.to(pending.dtype))

    def forward(self, img_z):
        """``x = attn(LN(x), q=x) + x; x = ff(LN(x)) + x`` per layer (reference ``:159-161``), scheduled so
        that every residual add is fused with the LayerNorm that follows it (``wm_add_layernorm_*``)."""
        return self._close(*self._blocks(img_z))

    def forward_last_frame(self, img_z, prune_receptive_field=False):
        """``forward(img_z)[:, -1]``: the only part of the output the denoiser head reads (``main.py:35``).  The final
        residual add (element-wise) is done on that frame alone -- same values, 1/S of the traffic.
        ``prune_receptive_field`` (opt-in): run every layer on ``last_frame_cone`` only, i.e. skip the rows whose
        contribution to that frame and to every parameter gradient is exactly zero (same values up to the summation
        order of the GEMMs; at depth 4, extents (1, ., .), 16 frames: 5, 4, 3, 2 frames instead of 16 per layer)."""
        x, pending, pending_bias = self._blocks(img_z, self.last_frame_cone(img_z.shape[1]) if prune_receptive_field else None)
        if pending is None:
            return x[:, -1]
        return ops.last_frame_close(x, pending, pending_bias)
