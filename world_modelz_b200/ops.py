"""Tensor-level entry points over the C ABI (``include/wm_b200.h``).

PyTorch is plumbing here: it owns device memory and the stream; all arithmetic of the
hot path happens inside ``libwm_b200.so``.  Every function raises if the tensor is not
on a CUDA device -- there is deliberately no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DTYPE_BF16, DTYPE_FP32, FLAG_SIMT, check

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
    _count(1)
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
    _count(2)
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


# ------------------------------------------------------------------------------------ VQ
def vq_nearest(x: torch.Tensor, codebook: torch.Tensor, want_quantized: bool = True,
               want_err: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
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
                                   N, L, K, D, DTYPE_FP32, 0, _stream()), 'wm_vq_nearest')
    _count(1)
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


def adamw_step(master, shadow, grad, exp_avg, exp_avg_sq, dyn, beta1, beta2, eps, weight_decay, grad_scale) -> None:
    """One fused AdamW launch over flat buffers (see ``wm_adamw_step`` in the header)."""
    _require_cuda(master, grad)
    check(_lib.lib().wm_adamw_step(master.data_ptr(), shadow.data_ptr() if shadow is not None else None,
                                   grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), master.numel(),
                                   dyn.data_ptr(), beta1, beta2, eps, weight_decay, grad_scale,
                                   _dtype_code(grad), _stream()), 'wm_adamw_step')
    _count(1)
