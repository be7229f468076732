"""Oracle (CPU restatement) of the local windowed 3D attention path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference builds the window by zero-padding K/V, taking three strided
``unfold`` views and materialising a ``window x`` copy
(``local_3d_attention.py:57-69,82-87``).  The oracle states the same arithmetic as
an index gather: every query token owns a table of ``Wn = prod(2*e+1)`` neighbour
token ids (row-major over the (i j k) window offsets, the reference's flatten order
``:86-87``) plus a validity flag for neighbours that fall outside the grid.  A
neighbour outside the grid is a zero key whose score is replaced by ``-1e9``
(``:92-94``) and therefore gets exactly zero softmax weight in fp32.

Everything is functional: parameters come in as a flat ``dict`` that uses the
reference's ``state_dict`` key names, so the same dict drives the reference module,
the oracle and the CUDA drop-in.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

MASK_FILL = -1e9  # local_3d_attention.py:93


# --------------------------------------------------------------------------- window
def window_table(S: int, H: int, W: int, extents: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Neighbour ids and validity for every token of an (S,H,W) grid.

    Returns ``ids [S*H*W, Wn]`` (int64, 0 where invalid) and ``valid [S*H*W, Wn]``.
    Offset order is row-major over (ds, dh, dw) in [-e, +e], which is the order the
    reference obtains from pad (``:57-63``) + unfold (``:65-69``) + ``(i j k)`` flatten.
    """
    eS, eH, eW = (int(e) for e in extents)
    s = torch.arange(S).view(S, 1, 1, 1, 1, 1)
    h = torch.arange(H).view(1, H, 1, 1, 1, 1)
    w = torch.arange(W).view(1, 1, W, 1, 1, 1)
    ds = torch.arange(-eS, eS + 1).view(1, 1, 1, -1, 1, 1)
    dh = torch.arange(-eH, eH + 1).view(1, 1, 1, 1, -1, 1)
    dw = torch.arange(-eW, eW + 1).view(1, 1, 1, 1, 1, -1)
    ks, kh, kw = s + ds, h + dh, w + dw
    valid = (ks >= 0) & (ks < S) & (kh >= 0) & (kh < H) & (kw >= 0) & (kw < W)
    ids = (ks.clamp(0, S - 1) * H + kh.clamp(0, H - 1)) * W + kw.clamp(0, W - 1)
    ids = torch.where(valid, ids, torch.zeros_like(ids))
    n = S * H * W
    return ids.reshape(n, -1), valid.reshape(n, -1)


def window_size(extents: Sequence[int]) -> int:
    return int(math.prod(2 * int(e) + 1 for e in extents))


# ------------------------------------------------------------------- attention core
def attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int,
                   extents: Sequence[int], scale: Optional[float] = None,
                   chunk: int = 1024, want_lse: bool = False):
    """``Local3dAttention.local_attention`` (``local_3d_attention.py:78-99``).

    ``q, k, v``: ``[B,S,H,W,heads*d]`` (the outputs of to_q / to_k / to_v), channels
    split head-major ``(H d)`` (``:85-87``).  Returns ``[B,S,H,W,heads*d]`` laid out
    the way ``forward`` re-merges heads (``:115``) and optionally the natural-log
    LSE ``[B,S,H,W,heads]``.  Differentiable (autograd); chunked over queries so that
    config-4 sized grids fit in host memory.
    """
    B, S, H, W, C = q.shape
    d = C // heads
    if scale is None:
        scale = d ** -0.5  # local_3d_attention.py:43
    N = S * H * W
    ids, valid = window_table(S, H, W, extents)
    eS, eH, eW = (int(e) for e in extents)
    Wn = window_size(extents)
    # Whole grid at once when the window-times copy of K and V fits comfortably (what the reference itself does,
    # local_3d_attention.py:82-87: zero pad, three unfold views, one materialising rearrange); otherwise the chunked
    # gather below.  Same arithmetic either way; the fast path keeps the CPU baseline as fast as the reference.
    if 2 * B * N * Wn * C * q.element_size() <= (6 << 30):
        def windows(t):                                # [B,S,H,W,C] -> [B, N, heads, d, Wn], (i j k) row-major
            tp = F.pad(t, (0, 0, eW, eW, eH, eH, eS, eS))
            tu = tp.unfold(1, 2 * eS + 1, 1).unfold(2, 2 * eH + 1, 1).unfold(3, 2 * eW + 1, 1)
            return tu.reshape(B, N, heads, d, Wn)
        dots = torch.einsum('bnhd,bnhdw->bnhw', q.reshape(B, N, heads, d), windows(k)) * scale
        dots = dots.masked_fill(~valid[None, :, None, :], MASK_FILL)
        out = torch.einsum('bnhw,bnhdw->bnhd', torch.softmax(dots, dim=-1), windows(v)).reshape(B, S, H, W, C)
        if want_lse:
            return out, torch.logsumexp(dots, dim=-1).reshape(B, S, H, W, heads)
        return out
    qf = q.reshape(B, N, heads, d)
    kf = k.reshape(B, N, heads, d)
    vf = v.reshape(B, N, heads, d)
    outs, lses = [], []
    for c0 in range(0, N, chunk):
        sel = ids[c0:c0 + chunk]                       # [n, Wn]
        ok = valid[c0:c0 + chunk]
        kg = kf[:, sel]                                # [B, n, Wn, heads, d]
        vg = vf[:, sel]
        dots = torch.einsum('bnhd,bnwhd->bnhw', qf[:, c0:c0 + chunk], kg) * scale
        dots = dots.masked_fill(~ok[None, :, None, :], MASK_FILL)
        if want_lse:
            lses.append(torch.logsumexp(dots, dim=-1))
        attn = torch.softmax(dots, dim=-1)
        outs.append(torch.einsum('bnhw,bnwhd->bnhd', attn, vg))
    out = torch.cat(outs, dim=1).reshape(B, S, H, W, C)
    if want_lse:
        return out, torch.cat(lses, dim=1).reshape(B, S, H, W, heads)
    return out


@torch.no_grad()
def attention_core_backward(q, k, v, dout, heads: int, extents: Sequence[int],
                            scale: Optional[float] = None, chunk: int = 1024):
    """Closed-form gradients of :func:`attention_core` (no autograd graph kept).

    With ``P = softmax(S)``, ``S = scale * q.k``:  ``dV[j] += P[i,j] dO[i]``,
    ``dP = dO.v``, ``dS = P * (dP - sum_j P dP)``, ``dQ[i] = scale * sum_j dS k_j``,
    ``dK[j] += scale * dS[i,j] q_i``.  Used for grids where keeping the autograd
    graph of the gathered K/V would not fit (config 4).
    """
    B, S, H, W, C = q.shape
    d = C // heads
    if scale is None:
        scale = d ** -0.5
    N = S * H * W
    ids, valid = window_table(S, H, W, extents)
    qf, kf, vf = (t.reshape(B, N, heads, d) for t in (q, k, v))
    dof = dout.reshape(B, N, heads, d)
    dq = torch.zeros_like(qf)
    dk = torch.zeros_like(kf)
    dv = torch.zeros_like(vf)
    for c0 in range(0, N, chunk):
        sel = ids[c0:c0 + chunk]
        ok = valid[c0:c0 + chunk]
        n, Wn = sel.shape
        kg, vg = kf[:, sel], vf[:, sel]
        qc, doc = qf[:, c0:c0 + chunk], dof[:, c0:c0 + chunk]
        dots = torch.einsum('bnhd,bnwhd->bnhw', qc, kg) * scale
        dots = dots.masked_fill(~ok[None, :, None, :], MASK_FILL)
        p = torch.softmax(dots, dim=-1)
        dp = torch.einsum('bnhd,bnwhd->bnhw', doc, vg)
        ds = p * (dp - (p * dp).sum(-1, keepdim=True)) * scale
        dq[:, c0:c0 + chunk] = torch.einsum('bnhw,bnwhd->bnhd', ds, kg)
        flat = sel.reshape(-1)
        dk_c = torch.einsum('bnhw,bnhd->bnwhd', ds, qc).reshape(B, n * Wn, heads, d)
        dv_c = torch.einsum('bnhw,bnhd->bnwhd', p, doc).reshape(B, n * Wn, heads, d)
        dk.index_add_(1, flat, dk_c)   # invalid slots carry exactly-zero weights
        dv.index_add_(1, flat, dv_c)
    shp = (B, S, H, W, C)
    return dq.reshape(shp), dk.reshape(shp), dv.reshape(shp)


@torch.no_grad()
def attention_core_slab(q, k, v, dout, heads: int, extents: Sequence[int], query_ids: torch.Tensor,
                        scale: Optional[float] = None, chunk: int = 512):
    """Forward and closed-form gradients of :func:`attention_core` restricted to the queries ``query_ids`` (flat
    token ids of the (S,H,W) grid).  Returns ``(out_q, dq_q, dk, dv)``: outputs and dQ for those queries
    ``[B, len(ids), C]``, and dK / dV ``[B,S,H,W,C]`` holding ONLY those queries' contributions -- complete for
    every key whose whole set of observers lies inside ``query_ids`` (e.g. the middle plane of a slab of
    ``2*eS+1`` query planes).  This is how full-size config 4 (32x32x32 tokens, window 5x7x7) is checked in
    seconds: same arithmetic as ``local_3d_attention.py:78-99``, a slab of queries at a time.
    """
    B, S, H, W, C = q.shape
    d = C // heads
    if scale is None:
        scale = d ** -0.5
    N = S * H * W
    ids, valid = window_table(S, H, W, extents)
    qf, kf, vf = (t.reshape(B, N, heads, d) for t in (q, k, v))
    dof = dout.reshape(B, N, heads, d)
    nq = query_ids.numel()
    out = torch.zeros(B, nq, heads, d)
    dq = torch.zeros(B, nq, heads, d)
    dk = torch.zeros_like(kf)
    dv = torch.zeros_like(vf)
    for c0 in range(0, nq, chunk):
        qi = query_ids[c0:c0 + chunk]
        sel, ok = ids[qi], valid[qi]
        n, Wn = sel.shape
        kg, vg = kf[:, sel], vf[:, sel]
        qc, doc = qf[:, qi], dof[:, qi]
        dots = torch.einsum('bnhd,bnwhd->bnhw', qc, kg) * scale
        dots = dots.masked_fill(~ok[None, :, None, :], MASK_FILL)
        p = torch.softmax(dots, dim=-1)
        out[:, c0:c0 + chunk] = torch.einsum('bnhw,bnwhd->bnhd', p, vg)
        dp = torch.einsum('bnhd,bnwhd->bnhw', doc, vg)
        ds = p * (dp - (p * dp).sum(-1, keepdim=True)) * scale
        dq[:, c0:c0 + chunk] = torch.einsum('bnhw,bnwhd->bnhd', ds, kg)
        flat = sel.reshape(-1)
        dk.index_add_(1, flat, torch.einsum('bnhw,bnhd->bnwhd', ds, qc).reshape(B, n * Wn, heads, d))
        dv.index_add_(1, flat, torch.einsum('bnhw,bnhd->bnwhd', p, doc).reshape(B, n * Wn, heads, d))
    return out.reshape(B, nq, C), dq.reshape(B, nq, C), dk.reshape(B, S, H, W, C), dv.reshape(B, S, H, W, C)


# ---------------------------------------------------------------------- module level
@dataclass
class DenoiserConfig:
    """Constructor arguments of ``VqVideoDiffusionModel`` (``main.py:25-31``)."""
    data_shape: Tuple[int, int, int]
    dim: int
    num_classes: int          # codebook size K; the input vocabulary is K+1 (mask token)
    extents: Tuple[int, int, int]
    depth: int
    heads: int
    dim_head: int
    mlp_dim: int

    @property
    def inner(self) -> int:
        return self.heads * self.dim_head

    @property
    def project_out(self) -> bool:  # local_3d_attention.py:40
        return not (self.heads == 1 and self.dim_head == self.dim)


def local3d_attention_module(p: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor,
                             q: torch.Tensor, heads: int, extents: Sequence[int]) -> torch.Tensor:
    """``Local3dAttention.forward(x, q)`` (``local_3d_attention.py:102-118``).

    K and V are projected from ``x``, Q from ``q`` (two different tensors under
    ``PreNorm``); ``to_v`` has a bias, ``to_q``/``to_k`` do not
    (``:46-48``); ``to_out`` is skipped when ``heads == 1 and dim_head == dim``.
    """
    kk = F.linear(x, p[prefix + 'to_k.weight'])
    vv = F.linear(x, p[prefix + 'to_v.weight'], p[prefix + 'to_v.bias'])
    qq = F.linear(q, p[prefix + 'to_q.weight'])
    core = attention_core(qq, kk, vv, heads, extents)
    if prefix + 'to_out.0.weight' in p:
        core = F.linear(core, p[prefix + 'to_out.0.weight'], p[prefix + 'to_out.0.bias'])
    return core.reshape(q.shape)


def positional_embedding(p: Dict[str, torch.Tensor], prefix: str, shape) -> torch.Tensor:
    """Sum of three axis embeddings (``local_3d_attention.py:140-151``)."""
    _, S, H, W = shape
    ps = p[prefix + 'pos_emb_s.weight'][:S].view(S, 1, 1, -1)
    ph = p[prefix + 'pos_emb_h.weight'][:H].view(1, H, 1, -1)
    pw = p[prefix + 'pos_emb_w.weight'][:W].view(1, 1, W, -1)
    return (ps + ph + pw).unsqueeze(0)


def transformer_forward(p: Dict[str, torch.Tensor], tokens: torch.Tensor, cfg: DenoiserConfig,
                        prefix: str = 'transformer.') -> torch.Tensor:
    """``Local3dAttentionTransformer.forward`` (``local_3d_attention.py:153-163``)."""
    x = F.embedding(tokens, p[prefix + 'embedding.weight'])
    x = x + positional_embedding(p, prefix, tokens.shape)
    dim = x.shape[-1]
    for layer in range(cfg.depth):
        a = f'{prefix}layers.{layer}.0.'
        f = f'{prefix}layers.{layer}.1.'
        # PreNorm normalises only the K/V source; q= is passed through un-normalised (:16-17,160)
        xn = F.layer_norm(x, (dim,), p[a + 'norm.weight'], p[a + 'norm.bias'])
        x = local3d_attention_module(p, a + 'fn.', xn, x, cfg.heads, cfg.extents) + x
        xn = F.layer_norm(x, (dim,), p[f + 'norm.weight'], p[f + 'norm.bias'])
        hdn = F.gelu(F.linear(xn, p[f + 'fn.net.0.weight'], p[f + 'fn.net.0.bias']))
        x = F.linear(hdn, p[f + 'fn.net.3.weight'], p[f + 'fn.net.3.bias']) + x
    return x


def denoiser_forward(p: Dict[str, torch.Tensor], tokens: torch.Tensor, cfg: DenoiserConfig) -> torch.Tensor:
    """``VqVideoDiffusionModel.forward`` (``main.py:33-36``): logits of the last frame."""
    x = transformer_forward(p, tokens, cfg)
    return F.linear(x[:, -1], p['logit_proj.weight'], p['logit_proj.bias'])


def init_denoiser_params(cfg: DenoiserConfig, seed: int = 42, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's shapes and key names (synthetic weights).

    The distributions mirror the stock ``nn.Embedding`` (N(0,1)) / ``nn.Linear``
    (U(+-1/sqrt(fan_in))) / ``nn.LayerNorm`` (1, 0) initialisers, drawn from one seeded
    generator so that every rank / implementation can rebuild identical weights.
    """
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound if bias else None
        return w, b

    p: Dict[str, torch.Tensor] = {}
    t = 'transformer.'
    p[t + 'embedding.weight'] = torch.randn(cfg.num_classes + 1, cfg.dim, generator=g)
    for name, n in zip(('pos_emb_s', 'pos_emb_h', 'pos_emb_w'), cfg.data_shape):
        p[f'{t}{name}.weight'] = torch.randn(n, cfg.dim, generator=g)
    for layer in range(cfg.depth):
        a = f'{t}layers.{layer}.0.'
        f = f'{t}layers.{layer}.1.'
        p[a + 'norm.weight'] = torch.ones(cfg.dim)
        p[a + 'norm.bias'] = torch.zeros(cfg.dim)
        p[a + 'fn.to_q.weight'], _ = lin(cfg.inner, cfg.dim, bias=False)
        p[a + 'fn.to_k.weight'], _ = lin(cfg.inner, cfg.dim, bias=False)
        p[a + 'fn.to_v.weight'], p[a + 'fn.to_v.bias'] = lin(cfg.inner, cfg.dim)
        if cfg.project_out:
            p[a + 'fn.to_out.0.weight'], p[a + 'fn.to_out.0.bias'] = lin(cfg.dim, cfg.inner)
        p[f + 'norm.weight'] = torch.ones(cfg.dim)
        p[f + 'norm.bias'] = torch.zeros(cfg.dim)
        p[f + 'fn.net.0.weight'], p[f + 'fn.net.0.bias'] = lin(cfg.mlp_dim, cfg.dim)
        p[f + 'fn.net.3.weight'], p[f + 'fn.net.3.bias'] = lin(cfg.dim, cfg.mlp_dim)
    p['logit_proj.weight'], p['logit_proj.bias'] = lin(cfg.num_classes, cfg.dim)
    return {k: v.to(dtype) for k, v in p.items()}


# ------------------------------------------------------------- training-step pieces
def corrupt_last_frame(tokens: torch.Tensor, r: torch.Tensor, num_embeddings: int,
                       gen: Optional[torch.Generator] = None, p_max_uniform: float = 0.1):
    """Noise + mask corruption of the last frame (``main.py:238-259``).

    ``tokens [B,S,H,W]`` int64, ``r [B]`` in [0,1).  With probability ``r*0.1`` a token
    is resampled uniformly (lerp of one-hot towards uniform, ``:251-255``), then every
    token is replaced by the mask token ``K`` with probability ``r`` (``:249,258``).
    Returns ``(corrupted tokens, target last frame)``.  RNG-dependent, so parity on
    this piece is distributional, not bitwise.
    """
    B = tokens.shape[0]
    last = tokens[:, -1]
    target = last.clone()
    enc = last.reshape(B, -1)
    rr = r.view(B, 1)
    masked = torch.rand(enc.shape, generator=gen) < rr
    probs = F.one_hot(enc, num_embeddings).float()
    probs = probs + (1.0 / num_embeddings - probs) * (rr.unsqueeze(-1) * p_max_uniform)
    draw = torch.multinomial(probs.view(-1, num_embeddings), 1, generator=gen).view(B, -1)
    draw = torch.where(masked, torch.full_like(draw, num_embeddings), draw)
    out = tokens.clone()
    out[:, -1] = draw.view(last.shape)
    return out, target


def denoiser_loss(p: Dict[str, torch.Tensor], tokens: torch.Tensor, target: torch.Tensor,
                  cfg: DenoiserConfig) -> torch.Tensor:
    """Mean cross-entropy of the last-frame logits (``main.py:266-274``)."""
    logits = denoiser_forward(p, tokens, cfg)
    return F.cross_entropy(logits.reshape(-1, cfg.num_classes), target.reshape(-1))


def train_step(p: Dict[str, torch.Tensor], opt_state: Dict[str, Dict[str, torch.Tensor]], step: int,
               tokens: torch.Tensor, target: torch.Tensor, cfg: DenoiserConfig,
               lr: float = 1e-4, weight_decay: float = 1e-7, betas=(0.9, 0.999), eps: float = 1e-8) -> float:
    """One optimisation step: forward, CE loss, backward, AdamW (``main.py:266-283,433``).

    Parameters are updated in place; ``opt_state[name] = {'m', 'v'}``; ``step`` is 1-based.
    """
    leaves = {k: v.detach().requires_grad_(True) for k, v in p.items()}
    loss = denoiser_loss(leaves, tokens, target, cfg)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    b1, b2 = betas
    with torch.no_grad():
        for (name, w), g in zip(p.items(), grads):
            st = opt_state.setdefault(name, {'m': torch.zeros_like(w), 'v': torch.zeros_like(w)})
            w.mul_(1 - lr * weight_decay)
            st['m'].mul_(b1).add_(g, alpha=1 - b1)
            st['v'].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (st['v'].sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
            w.addcdiv_(st['m'], denom, value=-lr / (1 - b1 ** step))
    return float(loss.detach())


def top_k_logits(logits: torch.Tensor, k: int) -> torch.Tensor:
    """Keep the ``k`` largest logits of every row, ``-inf`` elsewhere (``main.py:39-43``; ties at the k-th value stay)."""
    kth = torch.topk(logits, k, dim=-1).values[:, -1:]
    return logits.masked_fill(logits < kth, float('-inf'))


@torch.no_grad()
def sample_next_frame(p: Dict[str, torch.Tensor], tokens: torch.Tensor, cfg: DenoiserConfig,
                      iterations: int = 30, gen: Optional[torch.Generator] = None, sample_topk: int = -1) -> torch.Tensor:
    """Iterative mask/replace denoising of the last frame (``main.py:71-111``).

    ``tokens [B,S,H,W]``; the last frame's content is irrelevant (iteration 0 draws from flat logits and re-masks).
    Every iteration samples all positions from the current (optionally top-k filtered, ``:83-84``) logits, re-masks a
    ``1-(i+1)/iterations`` fraction and runs one denoiser forward.  ``tokens`` is updated IN PLACE like the
    reference's ``batch_z`` (``:107-109``); returns the final sampled last frame ``[B,H,W]``.
    """
    B, _, H, W = tokens.shape
    K = cfg.num_classes
    logits = torch.zeros(B * H * W, K)
    sample = None
    for i in range(iterations):
        if sample_topk > 0:
            logits = top_k_logits(logits, sample_topk)
        probs = torch.softmax(logits, dim=-1)
        sample = torch.multinomial(probs, 1, replacement=True, generator=gen).view(B, H, W)
        alpha = min(max((i + 1) / iterations, 0.0), 1.0)
        remask = torch.rand(B, H * W, generator=gen) > alpha
        tokens[:, -1] = torch.where(remask.view(B, H, W), torch.full_like(sample, K), sample)
        logits = denoiser_forward(p, tokens, cfg).reshape(B * H * W, K)
    return sample


@torch.no_grad()
def sample_frames(p: Dict[str, torch.Tensor], tokens: torch.Tensor, cfg: DenoiserConfig, num_steps: int,
                  iterations: int = 30, gen: Optional[torch.Generator] = None, sample_topk: int = -1) -> torch.Tensor:
    """The outer loop of ``evaluate_model`` (``main.py:71-115``): ``num_steps`` new frames, each by
    :func:`sample_next_frame`, with the context shifted by one frame in between (``:115``; the reference's
    in-place overlapping assignment ``batch_z[:,:-1] = batch_z[:,1:]`` raises on current PyTorch, the intended
    shift is restated with a copy).  Returns the sampled token frames ``[num_steps, B, H, W]``.
    """
    work = tokens.clone()
    work[:, -1] = cfg.num_classes                       # main.py:62: destroy all information in the last frame
    frames = []
    for _ in range(num_steps):
        frames.append(sample_next_frame(p, work, cfg, iterations, gen, sample_topk))
        work[:, :-1] = work[:, 1:].clone()
    return torch.stack(frames)
