"""Tensor-level entry points over the C ABI (``include/wm_b200.h``).

PyTorch is plumbing here: it owns device memory and the stream; all arithmetic of the
hot path happens inside ``libwm_b200.so``.  Every function raises if the tensor is not
on a CUDA device -- there is deliberately no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DTYPE_BF16, DTYPE_FP32, FLAG_SIMT, check  # noqa: F401

_launches = 0        # kernels launched by this library (claimed in bench.py's gpu_launches)


def launch_count() -> int:
    return _launches


def _count(n: int) -> None:
    global _launches
    _launches += n


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    if t.dtype == torch.float32:
        return DTYPE_FP32
    raise TypeError(f'local 3D attention supports float32 (exact) and bfloat16 (tensor cores), got {t.dtype}')


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError('world_modelz_b200 ops run on a CUDA device only (no CPU fallback); got a '
                               f'{t.device} tensor')


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def uses_tensor_cores(S, H, W, heads, dim_head, extents, dtype=torch.bfloat16) -> bool:
    code = DTYPE_BF16 if dtype == torch.bfloat16 else DTYPE_FP32
    return bool(_lib.lib().wm_l3d_attn_uses_tensor_cores(S, H, W, heads, dim_head, *[int(e) for e in extents], code))


def attn_forward(q, k, v, heads: int, extents: Sequence[int], scale: float, flags: int = 0):
    """Raw forward: returns ``(out, lse)``; no autograd.  ``q/k/v [B,S,H,W,heads*d]``."""
    _require_cuda(q, k, v)
    B, S, H, W, C = q.shape
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    out = torch.empty_like(q)
    lse = torch.empty(B, S, H, W, heads, device=q.device, dtype=torch.float32)
    check(_lib.lib().wm_l3d_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                     B, S, H, W, heads, C // heads, *[int(e) for e in extents], float(scale),
                                     _dtype_code(q), flags, _stream()), 'wm_l3d_attn_fwd')
    _count(2 if q.dtype == torch.bfloat16 and not (flags & FLAG_SIMT) else 1)     # tensor-core forward + its fix-up scan
    return out, lse


def attn_backward(q, k, v, out, lse, dout, heads: int, extents: Sequence[int], scale: float, flags: int = 0):
    """Raw backward: returns ``(dq, dk, dv)``."""
    B, S, H, W, C = q.shape
    dout = dout.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty_like(lse)
    check(_lib.lib().wm_l3d_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                     dout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), delta.data_ptr(),
                                     B, S, H, W, heads, C // heads, *[int(e) for e in extents], float(scale),
                                     _dtype_code(q), flags, _stream()), 'wm_l3d_attn_bwd')
    _count(3)
    return dq, dk, dv


class _Local3dAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, heads, extents, scale, flags):
        out, lse = attn_forward(q, k, v, heads, extents, scale, flags)
        ctx.save_for_backward(q.contiguous(), k.contiguous(), v.contiguous(), out, lse)
        ctx.cfg = (heads, tuple(int(e) for e in extents), float(scale), flags)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        heads, extents, scale, flags = ctx.cfg
        dq, dk, dv = attn_backward(q, k, v, out, lse, dout, heads, extents, scale, flags)
        return dq, dk, dv, None, None, None, None


def local3d_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, extents: Sequence[int],
                      scale: Optional[float] = None, flags: int = 0) -> torch.Tensor:
    """Fused local windowed 3D attention core (differentiable).

    Same contract as ``Local3dAttention.local_attention`` + head merge
    (``local_3d_attention.py:78-99,115``): ``q/k/v [B,S,H,W,heads*d]`` in, same shape out;
    only O and the LSE are kept for backward (no window-times copies, no checkpointing).
    """
    if scale is None:
        scale = (q.shape[-1] // heads) ** -0.5
    return _Local3dAttentionFn.apply(q, k, v, heads, tuple(extents), scale, flags)


class _Local3dAttentionKvFn(torch.autograd.Function):
    """Attention core over a MERGED key/value projection: ``kv [B,S,H,W,2*inner]`` is the output of one GEMM with
    ``cat(to_k.weight, to_v.weight)``; the kernels address its two channel halves in place (``wm_l3d_attn_*_ld``,
    token stride ``2*inner``), and backward writes dK | dV side by side into one buffer, which is then the single
    gradient operand of that GEMM's dgrad and wgrad -- no split / cat copies, no ``dxk + dxv`` add."""

    @staticmethod
    def forward(ctx, q, kv, heads, extents, scale, flags):
        _require_cuda(q, kv)
        B, S, H, W, C = q.shape
        if kv.shape[-1] != 2 * C or kv.shape[:-1] != q.shape[:-1]:
            raise ValueError(f'kv must be [B,S,H,W,2*inner] next to q {tuple(q.shape)}, got {tuple(kv.shape)}')
        q, kv = q.contiguous(), kv.contiguous()
        esz = q.element_size()
        out = torch.empty_like(q)
        lse = torch.empty(B, S, H, W, heads, device=q.device, dtype=torch.float32)
        check(_lib.lib().wm_l3d_attn_fwd_ld(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + C * esz, out.data_ptr(),
                                            lse.data_ptr(), 0, 2 * C, B, S, H, W, heads, C // heads,
                                            *[int(e) for e in extents], float(scale), _dtype_code(q), flags, _stream()),
              'wm_l3d_attn_fwd_ld')
        _count(2 if q.dtype == torch.bfloat16 and not (flags & FLAG_SIMT) else 1)
        ctx.save_for_backward(q, kv, out, lse)
        ctx.cfg = (heads, tuple(int(e) for e in extents), float(scale), flags)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, kv, out, lse = ctx.saved_tensors
        heads, extents, scale, flags = ctx.cfg
        B, S, H, W, C = q.shape
        esz = q.element_size()
        dout = dout.contiguous()
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        delta = torch.empty_like(lse)
        check(_lib.lib().wm_l3d_attn_bwd_ld(q.data_ptr(), kv.data_ptr(), kv.data_ptr() + C * esz, out.data_ptr(),
                                            lse.data_ptr(), dout.data_ptr(), dq.data_ptr(), dkv.data_ptr(),
                                            dkv.data_ptr() + C * esz, delta.data_ptr(), 0, 2 * C, B, S, H, W, heads,
                                            C // heads, *extents, scale, _dtype_code(q), flags, _stream()),
              'wm_l3d_attn_bwd_ld')
        _count(3)
        return dq, dkv, None, None, None, None


def local3d_attention_kv(q: torch.Tensor, kv: torch.Tensor, heads: int, extents: Sequence[int],
                         scale: Optional[float] = None, flags: int = 0) -> torch.Tensor:
    """``local3d_attention(q, kv[..., :inner], kv[..., inner:])`` without materialising the two halves
    (reference ``local_3d_attention.py:106-107`` as one projection)."""
    if scale is None:
        scale = (q.shape[-1] // heads) ** -0.5
    return _Local3dAttentionKvFn.apply(q, kv, heads, tuple(extents), scale, flags)


# ------------------------------------------------------------------------------------ VQ
def vq_nearest(x: torch.Tensor, codebook: torch.Tensor, want_quantized: bool = True,
               want_err: bool = True, flags: int = 0) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """``x [N,L,D]`` fp32, ``codebook [L,K,D]`` fp32 -> ``(idx int64 [N,L], ste [N,L,D], sq_err [N,L])``."""
    _require_cuda(x, codebook)
    if x.dtype != torch.float32 or codebook.dtype != torch.float32:
        raise TypeError('vq_nearest is the fp32 (bit-exact index) path; cast latents and codebook to float32')
    x = x.contiguous()
    codebook = codebook.contiguous()
    N, L, D = x.shape
    K = codebook.shape[1]
    idx = torch.empty(N, L, device=x.device, dtype=torch.int64)
    ste = torch.empty_like(x) if want_quantized else None
    err = torch.empty(N, L, device=x.device, dtype=torch.float32) if want_err else None
    check(_lib.lib().wm_vq_nearest(x.data_ptr(), codebook.data_ptr(), idx.data_ptr(),
                                   ste.data_ptr() if ste is not None else None,
                                   err.data_ptr() if err is not None else None,
                                   N, L, K, D, DTYPE_FP32, flags, _stream()), 'wm_vq_nearest')
    tensor_core = D == 64 and K % 64 == 0 and 64 <= K <= 512 and not (flags & FLAG_SIMT) and N >= 1
    _count(2 if tensor_core else 1)          # tcgen05 filter + the exact settlement of its undecided rows
    return idx, ste, err


def vq_distance(x: torch.Tensor, codebook: torch.Tensor, normalize: bool) -> torch.Tensor:
    _require_cuda(x, codebook)
    x = x.contiguous().float()
    codebook = codebook.contiguous().float()
    N, L, D = x.shape
    K = codebook.shape[1]
    dist = torch.empty(N, L, K, device=x.device, dtype=torch.float32)
    check(_lib.lib().wm_vq_distance(x.data_ptr(), codebook.data_ptr(), dist.data_ptr(), N, L, K, D, int(normalize),
                                    _stream()), 'wm_vq_distance')
    _count(1)
    return dist


def adamw_step(master, shadow, grad, exp_avg, exp_avg_sq, dyn, beta1, beta2, eps, weight_decay, grad_scale,
               grad_sq=None) -> None:
    """One fused AdamW launch over flat buffers (see ``wm_adamw_step_norm`` in the header); ``grad_sq`` (two device
    floats) also receives the squared gradient norm of the step."""
    _require_cuda(master, grad)
    check(_lib.lib().wm_adamw_step_norm(master.data_ptr(), shadow.data_ptr() if shadow is not None else None,
                                        grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), master.numel(),
                                        dyn.data_ptr(), beta1, beta2, eps, weight_decay, grad_scale,
                                        _dtype_code(grad), grad_sq.data_ptr() if grad_sq is not None else None,
                                        _stream()), 'wm_adamw_step_norm')
    _count(1)


def vq_stats(x, idx, sq_err, counts, dw=None, acc_err=None) -> None:
    """Accumulate the per-code statistics of one quantizer forward (``vq.py:35-46``) into ``counts [L,K]``,
    ``dw [L,K,D]`` and ``acc_err [L,K]`` -- one kernel, no one-hot."""
    _require_cuda(idx, counts)
    N, L = idx.shape
    K = counts.shape[-1]
    D = x.shape[-1] if x is not None else 1
    check(_lib.lib().wm_vq_stats(x.data_ptr() if x is not None else None, idx.data_ptr(),
                                 sq_err.data_ptr() if sq_err is not None else None, counts.data_ptr(),
                                 dw.data_ptr() if dw is not None else None,
                                 acc_err.data_ptr() if acc_err is not None else None, N, L, K, D, _stream()), 'wm_vq_stats')
    _count(1)


def vq_onehot(idx: torch.Tensor, K: int) -> torch.Tensor:
    """``one_hot(idx) -> float32 [..., K]`` written in one pass (``vq.py:39``)."""
    _require_cuda(idx)
    idx = idx.contiguous()
    out = torch.empty(*idx.shape, K, device=idx.device, dtype=torch.float32)
    if K % 4 != 0:
        return out.zero_().scatter_(-1, idx.unsqueeze(-1), 1.0)
    check(_lib.lib().wm_vq_onehot(idx.data_ptr(), out.data_ptr(), idx.numel(), K, _stream()), 'wm_vq_onehot')
    _count(1)
    return out


def sample_step(logits, sample, frame, frame_stride, per_clip, topk, mask_token, dyn, seed) -> None:
    """One draw of the iterative sampler for every position (``wm_sample_step``): ``logits [P,K]`` -> ``sample [P]``
    and the re-masked last frame written through ``frame`` (a view of the token tensor's last frame)."""
    _require_cuda(logits, sample)
    logits = logits.contiguous()
    P, K = logits.shape
    check(_lib.lib().wm_sample_step(logits.data_ptr(), sample.data_ptr(), frame.data_ptr() if frame is not None else None,
                                    frame_stride, per_clip, P, K, topk, mask_token, dyn.data_ptr(), seed,
                                    _dtype_code(logits), _stream()), 'wm_sample_step')
    _count(1)


def loss_hist_update(ts, losses, weights, counts, alpha) -> None:
    _require_cuda(ts, weights)
    check(_lib.lib().wm_loss_hist_update(ts.data_ptr(), losses.data_ptr(), weights.data_ptr(), counts.data_ptr(),
                                         ts.numel(), weights.numel(), float(alpha), _stream()), 'wm_loss_hist_update')
    _count(1)


class _EmbedPosFn(torch.autograd.Function):
    """``embedding(tokens) + ((pos_s + pos_h) + pos_w)`` in one kernel (``wm_embed_pos_fwd``); the backward is one
    one-hot GEMM (see ``backward``)."""

    @staticmethod
    def forward(ctx, tokens, table, ps, ph, pw):
        _require_cuda(tokens, table)
        B, S, H, W = tokens.shape
        dim = table.shape[1]
        tokens = tokens.contiguous()
        table, ps, ph, pw = (t.contiguous() for t in (table, ps, ph, pw))
        out = torch.empty(B, S, H, W, dim, device=table.device, dtype=table.dtype)
        check(_lib.lib().wm_embed_pos_fwd(tokens.data_ptr(), table.data_ptr(), ps.data_ptr(), ph.data_ptr(), pw.data_ptr(),
                                          out.data_ptr(), B, S, H, W, dim, table.shape[0], _dtype_code(table), _stream()),
              'wm_embed_pos_fwd')
        _count(1)
        ctx.save_for_backward(tokens)
        ctx.rows = (table.shape[0], ps.shape[0], ph.shape[0], pw.shape[0])
        return out

    @staticmethod
    def backward(ctx, dout):
        """All four gradients are ONE GEMM: ``onehot^T @ dout`` with a ``[tokens, rows + S + H + W]`` matrix that has four
        ones per row (the token's table row and its three position rows).  Products by 1.0 are exact and the GEMM
        accumulates in fp32, so this equals the stock embedding backward + broadcast reductions -- without the radix
        sort, the segment scan and the three full-size reductions (≈ 0.25 ms -> ≈ 0.08 ms at config 3)."""
        (tokens,) = ctx.saved_tensors
        n, ns, nh, nw = ctx.rows
        B, S, H, W = tokens.shape
        if ns != S or nh != H or nw != W:
            raise RuntimeError('embed_pos: position tables must be sliced to the grid (pos_s[:S], pos_h[:H], pos_w[:W])')
        dim = dout.shape[-1]
        ntok = tokens.numel()
        cols = (n + S + H + W + 7) // 8 * 8
        idx = torch.empty(ntok, 4, device=dout.device, dtype=torch.int64)
        idx[:, 0] = tokens.reshape(-1).clamp(0, n - 1)
        idx[:, 1:] = _pos_rows(B, S, H, W, n, dout.device)
        onehot = torch.zeros(ntok, cols, device=dout.device, dtype=dout.dtype)
        onehot.scatter_(1, idx, 1.0)
        g = onehot.t() @ dout.reshape(ntok, dim)
        return None, g[:n], g[n:n + S], g[n + S:n + S + H], g[n + S + H:n + S + H + W]


_pos_rows_cache = {}


def _pos_rows(B, S, H, W, n, device):
    """``[B*S*H*W, 3]`` int64: the one-hot columns of a token's three position rows (constant per shape, cached)."""
    key = (B, S, H, W, n, str(device))
    t = _pos_rows_cache.get(key)
    if t is None:
        s = torch.arange(S, device=device).view(S, 1, 1).expand(S, H, W)
        h = torch.arange(H, device=device).view(1, H, 1).expand(S, H, W)
        w = torch.arange(W, device=device).view(1, 1, W).expand(S, H, W)
        one = torch.stack((n + s, n + S + h, n + S + H + w), dim=-1).reshape(1, S * H * W, 3)
        t = one.expand(B, -1, -1).reshape(B * S * H * W, 3).contiguous()
        if len(_pos_rows_cache) > 16:
            _pos_rows_cache.clear()
        _pos_rows_cache[key] = t
    return t


def embed_pos(tokens: torch.Tensor, table: torch.Tensor, pos_s: torch.Tensor, pos_h: torch.Tensor,
              pos_w: torch.Tensor) -> torch.Tensor:
    """Token embedding + the three axis position embeddings (``local_3d_attention.py:143-157``) -> ``[B,S,H,W,dim]``."""
    _require_cuda(tokens, table)
    if table.shape[1] % 8 != 0 or table.dtype not in (torch.bfloat16, torch.float32):
        _, s, h, w = tokens.shape
        pos = pos_s[:s, None, None, :] + pos_h[None, :h, None, :] + pos_w[None, None, :w, :]
        return torch.nn.functional.embedding(tokens, table) + pos.unsqueeze(0)
    return _EmbedPosFn.apply(tokens, table, pos_s, pos_h, pos_w)


def _colsum(t2d: torch.Tensor) -> torch.Tensor:
    """Column sums of a contiguous ``[rows, cols]`` tensor with ``wm_colsum`` (stock reductions of a tall, narrow matrix
    run on a handful of CTAs: ~20 us for 8192 x 256)."""
    rows, cols = t2d.shape
    out = torch.empty(cols, device=t2d.device, dtype=t2d.dtype)
    ws = torch.empty(_lib.lib().wm_reduce_blocks(rows) * cols, device=t2d.device, dtype=torch.float32)
    check(_lib.lib().wm_colsum(t2d.data_ptr(), out.data_ptr(), ws.data_ptr(), rows, cols, _dtype_code(t2d), _stream()),
          'wm_colsum')
    _count(2)
    return out


class _LastFrameCloseFn(torch.autograd.Function):
    """``x[:, -1] + (pending[:, -1] + bias)``: the last residual add of the transformer, on the only frame the denoiser
    head reads (``main.py:35``).  Backward: the gradients of ``x`` and ``pending`` are the SAME tensor (zero but for the last
    frame) -- one fill and one copy instead of two of each -- and the bias gradient is a ``wm_colsum``."""

    @staticmethod
    def forward(ctx, x, pending, bias):
        ctx.shape = x.shape
        ctx.has_bias = bias is not None
        last = pending[:, -1] if bias is None else pending[:, -1] + bias.to(pending.dtype)
        return x[:, -1] + last

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        g = torch.zeros(ctx.shape, device=dout.device, dtype=dout.dtype)
        g[:, -1] = dout
        dbias = None
        if ctx.has_bias:
            cols = dout.shape[-1]
            d2 = dout.reshape(-1, cols)
            dbias = _colsum(d2) if (cols % 8 == 0 and cols <= 2048) else d2.sum(0)
        return g, g, dbias


def last_frame_close(x: torch.Tensor, pending: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    _require_cuda(x, pending)
    return _LastFrameCloseFn.apply(x, pending, bias)


class _ProjPassFn(torch.autograd.Function):
    """``(x @ W^T, x)``: a bias-free projection whose input also continues down another branch (``to_q`` reads the
    residual stream, ``local_3d_attention.py:160``).  Routing the stream THROUGH this node lets the backward fold the
    stream's gradient into the projection's dgrad GEMM (``beta = 1``) instead of a separate full-size add."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return torch.nn.functional.linear(x, weight), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dpass):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dw = dy2.t() @ x2
        if dpass is None:
            dx = (dy2 @ weight).view(x.shape)
        else:
            dx = torch.addmm(dpass.reshape(-1, x.shape[-1]), dy2, weight).view(x.shape)
        return dx, dw


def linear_passthrough(x: torch.Tensor, weight: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _require_cuda(x)
    return _ProjPassFn.apply(x, weight)


# ------------------------------------------------------- layer-level kernels (LayerNorm, bias grads)
def _rows(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


class _AddLayerNormFn(torch.autograd.Function):
    """(res, delta, delta_bias) -> (sum = res + delta + delta_bias, LayerNorm(sum)); ``delta`` may be None (then sum is
    res) and so may ``delta_bias`` (the [dim] bias of the linear layer that produced ``delta``)."""

    @staticmethod
    def forward(ctx, res, delta, delta_bias, gamma, beta, eps):
        _require_cuda(res)
        res = res.contiguous()
        dim = res.shape[-1]
        rows = _rows(res)
        y = torch.empty_like(res)
        mean = torch.empty(rows, device=res.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=res.device, dtype=torch.float32)
        if delta is not None:
            delta = delta.contiguous()
            total = torch.empty_like(res)
            if delta_bias is not None:
                delta_bias = delta_bias.to(res.dtype).contiguous()
        else:
            if delta_bias is not None:
                raise ValueError('delta_bias without delta')
            total = res
        check(_lib.lib().wm_add_layernorm_fwd(res.data_ptr(), delta.data_ptr() if delta is not None else None,
                                              delta_bias.data_ptr() if delta_bias is not None else None,
                                              gamma.data_ptr(), beta.data_ptr(),
                                              total.data_ptr() if delta is not None else None, y.data_ptr(),
                                              mean.data_ptr(), rstd.data_ptr(), rows, dim, float(eps), _dtype_code(res),
                                              _stream()), 'wm_add_layernorm_fwd')
        _count(1)
        ctx.save_for_backward(total, mean, rstd, gamma)
        ctx.has_delta = delta is not None
        ctx.has_bias = delta_bias is not None
        return total, y

    @staticmethod
    def backward(ctx, dtotal, dy):
        total, mean, rstd, gamma = ctx.saved_tensors
        dim = total.shape[-1]
        rows = _rows(total)
        dy = dy.contiguous() if dy is not None else torch.zeros_like(total)
        dres = dtotal.contiguous() if dtotal is not None else None
        dx = torch.empty_like(total)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        dbias = torch.empty_like(gamma) if ctx.has_bias else None
        ws = torch.empty(_lib.lib().wm_reduce_blocks(rows) * 3 * dim, device=total.device, dtype=torch.float32)
        check(_lib.lib().wm_add_layernorm_bwd(dy.data_ptr(), dres.data_ptr() if dres is not None else None,
                                              total.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                              dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                              dbias.data_ptr() if dbias is not None else None, ws.data_ptr(), rows,
                                              dim, _dtype_code(total), _stream()), 'wm_add_layernorm_bwd')
        _count(2)
        return dx, (dx if ctx.has_delta else None), dbias, dgamma, dbeta, None


def add_layernorm(res: torch.Tensor, delta: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float = 1e-5, delta_bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused ``s = res + delta (+ delta_bias); y = LayerNorm(s)``; returns ``(s, y)`` (``s is res`` when delta is None).

    One kernel forward, one backward (which also folds the gradient arriving at ``s`` from the residual path into
    ``dx`` and, when ``delta_bias`` is given, reduces the bias gradient in the same pass).  ``delta_bias`` is the
    bias of the linear layer that produced ``delta`` (``to_out.0`` / ``net.3``), deferred to here.
    Reference: PreNorm + residual adds, local_3d_attention.py:11-17,159-161.
    """
    _require_cuda(res)
    dim = res.shape[-1]
    if dim % 8 != 0 or dim > 2048:      # widths the kernel does not tile: stock CUDA ops (still on the device)
        total = res if delta is None else (res + delta if delta_bias is None else res + (delta + delta_bias))
        return total, torch.nn.functional.layer_norm(total, (dim,), gamma, beta, eps)
    return _AddLayerNormFn.apply(res, delta, delta_bias, gamma, beta, eps)


class _LinearFn(torch.autograd.Function):
    """``x @ W^T + b`` on cuBLAS; the bias gradient is a column sum in ``wm_colsum``."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = x.reshape(-1, x.shape[-1])
        dx = (dy2 @ weight).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = dy2.t() @ x2
        db = None
        if ctx.has_bias:
            rows, cols = dy2.shape
            db = torch.empty(cols, device=dy.device, dtype=dy.dtype)
            ws = torch.empty(_lib.lib().wm_reduce_blocks(rows) * cols, device=dy.device, dtype=torch.float32)
            check(_lib.lib().wm_colsum(dy2.data_ptr(), db.data_ptr(), ws.data_ptr(), rows, cols, _dtype_code(dy2),
                                       _stream()), 'wm_colsum')
            _count(2)
        return dx, dw, db


class _BiasGeluFn(torch.autograd.Function):
    """``gelu(h + bias)`` (exact erf form) in one kernel; the backward writes ``dh`` and reduces ``dbias`` in one pass."""

    @staticmethod
    def forward(ctx, h, bias):
        _require_cuda(h)
        h = h.contiguous()
        bias = bias.to(h.dtype).contiguous()
        y = torch.empty_like(h)
        check(_lib.lib().wm_bias_gelu_fwd(h.data_ptr(), bias.data_ptr(), y.data_ptr(), _rows(h), h.shape[-1],
                                          _dtype_code(h), _stream()), 'wm_bias_gelu_fwd')
        _count(1)
        ctx.save_for_backward(h, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        h, bias = ctx.saved_tensors
        dy = dy.contiguous()
        rows, cols = _rows(h), h.shape[-1]
        dh = torch.empty_like(h)
        dbias = torch.empty_like(bias)
        ws = torch.empty(_lib.lib().wm_reduce_blocks(rows) * cols, device=h.device, dtype=torch.float32)
        check(_lib.lib().wm_bias_gelu_bwd(dy.data_ptr(), h.data_ptr(), bias.data_ptr(), dh.data_ptr(), dbias.data_ptr(),
                                          ws.data_ptr(), rows, cols, _dtype_code(h), _stream()), 'wm_bias_gelu_bwd')
        _count(2)
        return dh, dbias


def bias_gelu(h: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """``gelu(h + bias)`` with ``h = x @ W1^T`` computed WITHOUT bias (FeedForward ``net.0`` / ``net.1``,
    local_3d_attention.py:24-27).  Widths the kernel does not tile fall back to stock CUDA ops."""
    _require_cuda(h)
    cols = h.shape[-1]
    if cols % 8 != 0 or cols > 2048 or h.dtype not in (torch.bfloat16, torch.float32):
        return torch.nn.functional.gelu(h + bias)
    return _BiasGeluFn.apply(h, bias)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear forward on cuBLAS with the bias gradient reduced by ``wm_colsum``."""
    _require_cuda(x)
    if bias is None or bias.shape[0] % 8 != 0 or bias.shape[0] > 2048:
        return torch.nn.functional.linear(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)
