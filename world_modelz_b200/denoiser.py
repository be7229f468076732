"""Denoiser wrapper, data-parallel training step and iterative sampler.

* ``VqVideoDiffusionModel`` -- drop-in for ``vq-video-diffusion/main.py:25-36`` (same
  constructor keywords, ``transformer`` / ``logit_proj`` sub-modules and state_dict keys).
* ``corrupt_last_frame`` -- the noise + mask corruption of ``main.py:246-259`` on the
  device, without the ``[B, HW, K]`` one-hot / uniform temporaries.
* ``DenoiserTrainer`` -- one optimisation step (forward, CE, backward, gradient
  all-reduce, AdamW) as in ``main.py:266-283``; batch-sharded data parallel, one process
  per GPU, one NCCL all-reduce of a flat gradient buffer per step; forward+backward and
  the optimiser update are two CUDA graphs with the (eager) all-reduce between them.  The reference has no distributed code.
* ``sample_next_frame`` -- the 30-iteration mask/replace sampler of ``main.py:71-111``;
  clips are independent, so multi-GPU sampling is batch-sharded with no collective.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import ops
from .local_3d_attention import Local3dAttentionTransformer


class VqVideoDiffusionModel(nn.Module):
    def __init__(self, *, data_shape, dim, num_classes, extents, depth, dim_head, mlp_dim, heads=1, dropout=.0):
        super().__init__()
        self.num_classes = num_classes
        self.transformer = Local3dAttentionTransformer(
            data_shape=data_shape, dim=dim, num_classes=num_classes + 1,      # +1: the mask token
            extents=extents, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim, dropout=dropout)
        self.logit_proj = nn.Linear(dim, num_classes)
        # opt-in (not part of the reference's surface): evaluate only the rows the last frame depends on -- see
        # Local3dAttentionTransformer.last_frame_cone.  Off by default: every number quoted as the reference's metric
        # is measured on the full 16-frame computation.
        self.prune_receptive_field = False

    def forward(self, x):
        feats = self.transformer.forward_last_frame(x, self.prune_receptive_field)   # == transformer(x)[:, -1]
        return ops.linear(feats, self.logit_proj.weight, self.logit_proj.bias)   # nn.Linear; bias gradient by wm_colsum


def corrupt_last_frame(tokens: torch.Tensor, r: torch.Tensor, num_embeddings: int,
                       p_max_uniform: float = 0.1) -> Tuple[torch.Tensor, torch.Tensor]:
    """``tokens [B,S,H,W]`` int64 on the device, ``r [B]`` in [0,1) -> (corrupted, target).

    The reference draws from ``lerp(onehot(z), 1/K, r*0.1)`` with ``multinomial``
    (``main.py:251-255``); that distribution is "keep z with probability 1-r*0.1, else a
    uniform code", sampled here directly.  Then each position becomes the mask token
    ``K`` with probability ``r`` (``:249,258``).
    """
    B = tokens.shape[0]
    last = tokens[:, -1]
    target = last.clone()
    rr = r.view(B, 1, 1).to(torch.float32)
    u = torch.rand(3, *last.shape, device=tokens.device)
    uniform_code = (u[0] * num_embeddings).long().clamp_(max=num_embeddings - 1)
    draw = torch.where(u[1] < rr * p_max_uniform, uniform_code, last)
    draw = torch.where(u[2] < rr, torch.full_like(draw, num_embeddings), draw)
    out = tokens.clone()
    out[:, -1] = draw
    return out, target


class LossAwareSamplerEma:
    """Loss-aware diffusion-time sampler (``importance_sampling.py:5-47``): a 100-bucket EMA histogram of per-sample
    losses.  With ``device=None`` it lives on the host like the reference's (loss feedback may be handed over
    asynchronously, one step late, so that it never stalls the GPU).  With a CUDA ``device`` the histogram lives on
    the GPU: ``update_with_losses`` is one ``wm_loss_hist_update`` launch (same sequential-EMA order within a bucket,
    ``:40-41``) and ``sample`` draws with device ops -- the training loop then has no device-to-host copy at all.
    """

    def __init__(self, num_histogram_buckets=100, uniform_p=0.01, alpha=0.9, warmup=10, jitter=True, seed=0, device=None):
        self.n = num_histogram_buckets
        self.uniform_p, self.alpha, self.warmup, self.jitter = uniform_p, alpha, warmup, jitter
        self.device = torch.device(device) if device is not None else torch.device('cpu')
        self._weights = torch.ones(self.n, device=self.device)
        self._counts = torch.zeros(self.n, dtype=torch.long, device=self.device)
        self._gen = torch.Generator(device=self.device).manual_seed(seed)
        self._pending = None

    def warmed_up(self) -> bool:
        return bool((self._counts > self.warmup).all())

    def weights(self) -> torch.Tensor:
        if self.device.type == 'cuda':           # no host decision: blend by the (device) warm-up flag
            warm = (self._counts > self.warmup).all().float()
            w = self._weights / self._weights.sum()
            w = (1 - self.uniform_p) * w + self.uniform_p / self.n
            return warm * w + (1 - warm) * torch.ones_like(w)
        if not self.warmed_up():
            return torch.ones(self.n)
        w = self._weights / self._weights.sum()
        return (1 - self.uniform_p) * w + self.uniform_p / self.n

    def sample(self, batch_size: int) -> torch.Tensor:
        self._drain()
        b = torch.multinomial(self.weights(), batch_size, replacement=True, generator=self._gen).float()
        if self.jitter:
            return (b + torch.rand(batch_size, generator=self._gen, device=self.device)) / self.n
        return b / (self.n - 1)

    def update_with_losses(self, ts: torch.Tensor, losses: torch.Tensor) -> None:
        if self.device.type == 'cuda':
            ops.loss_hist_update(ts.reshape(-1).float().contiguous(), losses.reshape(-1).float().contiguous(),
                                 self._weights, self._counts, self.alpha)
            return
        idx = (ts.view(-1) * self.n).long().clamp_(0, self.n - 1)
        self._counts.scatter_add_(0, idx, torch.ones_like(idx))
        for i, j in enumerate(idx.tolist()):
            self._weights[j] = self._weights[j] * self.alpha + float(losses[i]) * (1 - self.alpha)

    def update_async(self, ts_host: torch.Tensor, losses_pinned: torch.Tensor, ready: torch.cuda.Event) -> None:
        self._drain()
        self._pending = (ts_host, losses_pinned, ready)

    def _drain(self) -> None:
        if self._pending is not None:
            ts, losses, ev = self._pending
            ev.synchronize()
            self.update_with_losses(ts, losses)
            self._pending = None


def _align(n: int, to: int = 8) -> int:
    return (n + to - 1) // to * to


class DenoiserTrainer:
    """One training step of the denoiser, batch-sharded over ``world_size`` GPUs.

    Parameters live in flat buffers: fp32 master weights, a compute-dtype shadow the modules read (``param.data`` are
    views into it) and an fp32 gradient buffer that one multi-tensor copy fills from the per-parameter gradients
    after backward.  Every parameter starts at a multiple of 8 elements (16 bytes in bf16) so that the kernels'
    alignment rule holds for any parameter shape; the padding carries zero gradients and stays zero.  A step is:
    zero grads, corruption, forward, mean CE, backward, ONE all-reduce(SUM, fp32) of the flat gradients over NCCL, ONE
    fused AdamW launch (``wm_adamw_step_norm``) that also refreshes the shadow copy and reduces the gradient norm.

    ``accumulation_steps`` (``main.py:158,205,221-280``): gradients of that many ``step`` calls are averaged before the
    exchange / update.  ``lr_schedule`` (``main.py:441-442``): optional ``f(optimizer_step) -> lr`` applied before
    every update.
    """

    def __init__(self, model: VqVideoDiffusionModel, *, lr=1e-4, weight_decay=1e-7, betas=(0.9, 0.999), eps=1e-8,
                 compute_dtype=torch.bfloat16, process_group=None, use_cuda_graph=True, accumulation_steps=1,
                 lr_schedule=None):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError('DenoiserTrainer needs the model on a CUDA device')
        self.model = model
        self.device = params[0].device
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.accumulation_steps = int(accumulation_steps)
        self.lr_schedule = lr_schedule
        self._micro = 0                       # micro-batches accumulated since the last update
        self._opt_steps = 0
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off = _align(off + p.numel())
        n = off
        self.n_params = sum(p.numel() for p in params)
        self.master = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.shadow = torch.zeros(n, device=self.device, dtype=compute_dtype) if compute_dtype != torch.float32 else None
        self.grad = torch.zeros(n, device=self.device, dtype=torch.float32)       # reduced across ranks in fp32
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.grad_sq = torch.zeros(2, device=self.device, dtype=torch.float32)    # squared gradient norm, by step parity
        self._params = params
        self._grad_views = []
        for p, off in zip(params, self.offsets):
            k = p.numel()
            self.master[off:off + k].copy_(p.detach().reshape(-1).float())
            store = self.shadow if self.shadow is not None else self.master
            store[off:off + k].copy_(p.detach().reshape(-1))
            p.data = store[off:off + k].view_as(p)
            self._grad_views.append(self.grad[off:off + k].view_as(p))
        if self.world > 1:   # identical replicas: rank 0's weights win
            dist.broadcast(self.master, src=dist.get_global_rank(self.pg, 0) if self.pg else 0, group=self.pg)
            if self.shadow is not None:
                self.shadow.copy_(self.master)
        self.dyn = torch.tensor([1.0, lr], device=self.device, dtype=torch.float32)      # {step, lr}
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static = None
        self.K = model.num_classes

    # -- the step, in three pieces: graph A | one NCCL all-reduce | graph B ------------------------
    def _forward_backward(self, tokens, r, keep):
        """``keep``: 0-d device tensor, 0.0 on the first micro-batch of an accumulation window (the flat gradient
        buffer is overwritten), 1.0 on the later ones (it is added to)."""
        for p in self._params:                # autograd then ASSIGNS fresh gradients instead of launching one add per parameter
            p.grad = None
        corrupted, target = corrupt_last_frame(tokens, r, self.K)
        logits = self.model(corrupted)
        ce = F.cross_entropy(logits.reshape(-1, self.K).float(), target.reshape(-1), reduction='none')
        per_sample = ce.view(tokens.shape[0], -1).mean(dim=1)
        loss = ce.mean()
        (loss / self.accumulation_steps if self.accumulation_steps > 1 else loss).backward()     # main.py:276-278
        grads = [p.grad for p in self._params]
        if self.accumulation_steps > 1:
            self.grad.mul_(keep)
            torch._foreach_add_(self._grad_views, [g.float() for g in grads])
        else:   # one multi-tensor copy gathers every parameter gradient into the flat buffer (the all-reduce / AdamW operand)
            torch._foreach_copy_(self._grad_views, grads)
        return loss.detach(), per_sample.detach()

    def _exchange(self):
        """The path's only collective: SUM all-reduce of the flat fp32 gradient buffer (NCCL over NVLink)."""
        if self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg)

    def _update(self):
        ops.adamw_step(self.master, self.shadow, self.grad, self.exp_avg, self.exp_avg_sq, self.dyn, self.betas[0],
                       self.betas[1], self.eps, self.weight_decay, 1.0 / self.world, self.grad_sq)
        self.dyn[0:1] += 1.0

    def set_lr(self, lr: float) -> None:
        self.dyn[1:2].fill_(lr)

    def grad_norm(self) -> torch.Tensor:
        """L2 norm of the (averaged) gradient of the last update, as a 0-d DEVICE tensor -- ``grad_norm(model)`` of
        ``main.py:189-193,282`` without its per-parameter ``.item()`` syncs (reduced inside the AdamW kernel)."""
        return self.grad_sq[(self._opt_steps & 1)].sqrt()

    def step(self, tokens: torch.Tensor, r: torch.Tensor):
        """``tokens [b,S,H,W]`` int64 and ``r [b]`` on the device -> (loss, per-sample loss) tensors.  With
        ``accumulation_steps == k`` every k-th call exchanges gradients and updates the weights."""
        first = self._micro == 0
        last = self._micro + 1 == self.accumulation_steps
        if self.use_cuda_graph:
            if self._graph is None:
                self._capture(tokens, r)
            st_tokens, st_r, st_keep, st_loss, st_ps = self._static
            st_tokens.copy_(tokens, non_blocking=True)
            st_r.copy_(r, non_blocking=True)
            if self.accumulation_steps > 1:
                st_keep.fill_(0.0 if first else 1.0)
            graph_a, graph_b = self._graph
            graph_a.replay()
            out = (st_loss, st_ps)
        else:
            keep = torch.full((), 0.0 if first else 1.0, device=self.device)
            out = self._forward_backward(tokens, r.float(), keep)
        self._micro += 1
        if last:
            self._micro = 0
            if self.lr_schedule is not None:
                self.set_lr(float(self.lr_schedule(self._opt_steps)))
            self._exchange()          # eager, between the two graphs: the collective is never captured
            if self.use_cuda_graph:
                self._graph[1].replay()
            else:
                self._update()
            self._opt_steps += 1
        return out

    def _capture(self, tokens, r):
        st_tokens, st_r = tokens.clone(), r.clone().float()
        st_keep = torch.zeros((), device=self.device)
        saved = (self.master.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.dyn.clone(),
                 None if self.shadow is None else self.shadow.clone(), self.grad_sq.clone())
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                 # warm-up on a side stream (allocator, cuBLAS handles, autograd)
                before = ops.launch_count()
                self._forward_backward(st_tokens, st_r, st_keep)
                self._update()
                self._launches_per_step = ops.launch_count() - before
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_a):
            st_loss, st_ps = self._forward_backward(st_tokens, st_r, st_keep)
        graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_b, pool=graph_a.pool()):
            self._update()
        # undo the warm-up updates so that step 1 is step 1 (capture itself executes nothing)
        self.master.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2])
        self.dyn.copy_(saved[3]); self.grad_sq.copy_(saved[5])
        if self.shadow is not None:
            self.shadow.copy_(saved[4])
        self._graph, self._static = (graph_a, graph_b), (st_tokens, st_r, st_keep, st_loss, st_ps)

    # -- checkpoints in the reference's format (main.py:297-309, 370-410) ------------------------------
    def checkpoint(self, step: int, opt=None, ema_model_state=None) -> dict:
        """The dict the reference's training loop saves: fp32 ``model_state_dict`` from the master weights, an
        ``optimizer_state_dict`` laid out like ``torch.optim.AdamW.state_dict()``, and ``ema_model_state_dict`` passed
        through when the caller keeps an EMA model (``main.py:303``)."""
        from . import checkpoint as ck
        self._sync_params_from_master()
        taken = int(round(float(self.dyn[0].item()))) - 1
        shapes = ck.param_shapes(self._params)
        opt_state = ck.flat_to_adamw_state(shapes, self.exp_avg, self.exp_avg_sq, taken, lr=float(self.dyn[1].item()),
                                           betas=self.betas, eps=self.eps, weight_decay=self.weight_decay,
                                           offsets=self.offsets)
        state = self.model.state_dict()
        named = {id(q): name for name, q in self.model.named_parameters()}
        for p, off in zip(self._params, self.offsets):        # parameters: full-precision master copy, not the bf16 shadow
            state[named[id(p)]] = self.master[off:off + p.numel()].view_as(p).clone()
        return ck.make_checkpoint(state, opt_state, step, float(self.dyn[1].item()), opt, ema_model_state)

    def load_checkpoint(self, data: dict) -> int:
        """Resume from a dict written by the reference (or by :meth:`checkpoint`); returns the stored step.  The
        reference itself restores only the weights (``main.py:367-372,409-410``); optimizer moments, step counter and
        learning rate are restored here as well when present."""
        from . import checkpoint as ck
        self.model.load_state_dict(data['model_state_dict'], strict=True)      # copies into the shadow views
        named = {id(q): name for name, q in self.model.named_parameters()}
        sd = data['model_state_dict']
        for p, off in zip(self._params, self.offsets):
            self.master[off:off + p.numel()].copy_(sd[named[id(p)]].reshape(-1).float())
        if data.get('optimizer_state_dict') is not None:
            taken = ck.adamw_state_to_flat(data['optimizer_state_dict'], ck.param_shapes(self._params), self.exp_avg,
                                           self.exp_avg_sq, offsets=self.offsets)
            self.dyn[0:1].fill_(float(taken + 1))
            self._opt_steps = taken
            self.grad_sq.zero_()
        lr = data.get('lr')
        if lr is not None:
            self.set_lr(float(lr[0] if isinstance(lr, (list, tuple)) else lr))
        return int(data.get('step', 0))

    def _sync_params_from_master(self) -> None:
        if self.shadow is not None:
            self.shadow.copy_(self.master)

    def launches_per_step(self) -> int:
        """Launches of libwm_b200 kernels per step (attention fwd/bwd, add+LayerNorm fwd/bwd, bias column sums,
        AdamW), counted by ``ops`` during the last eager warm-up step; the graphs replay exactly those."""
        return getattr(self, '_launches_per_step', 0)


class _SamplerState:
    """Buffers of one (model, shape) sampling loop: the working token tensor, the logits the next draw uses, the draw,
    ``dyn = {alpha, call counter}`` on the device, its per-iteration table and (optionally) the captured iteration."""

    def __init__(self, model, shape, device, iterations, seed):
        B, S, H, W = shape
        K = model.num_classes
        self.work = torch.zeros(B, S, H, W, dtype=torch.int64, device=device)
        self.logits = torch.zeros(B * H * W, K, device=device)
        self.sample = torch.zeros(B * H * W, dtype=torch.int64, device=device)
        self.dyn = torch.zeros(2, device=device)
        self.seed = seed
        self.calls = 0
        self.iterations = iterations
        self.alphas = torch.tensor([min(max((i + 1) / iterations, 0.0), 1.0) for i in range(iterations)], device=device)
        self.graph = None
        self.graph_topk = None

    def iteration(self, model, topk):
        """One mask/replace iteration (``main.py:79-111``): ``wm_sample_step`` draws every position from the current
        logits (top-k filtered when asked) and writes the re-masked last frame, then one denoiser forward."""
        B, S, H, W = self.work.shape
        last = self.work[:, -1]
        ops.sample_step(self.logits, self.sample, last, S * H * W, H * W, topk, model.num_classes, self.dyn, self.seed)
        self.logits.copy_(model(self.work).reshape(B * H * W, -1))


def _sampler_state(model, tokens, iterations, seed):
    cache = model.__dict__.setdefault('_wm_sample_graphs', {})
    key = (tuple(tokens.shape), str(tokens.device), model.training, iterations,
           bool(getattr(model, 'prune_receptive_field', False)))
    if key not in cache:
        cache[key] = _SamplerState(model, tokens.shape, tokens.device, iterations, seed)
    return cache[key]


def _denoise_last_frame(model, st, sample_topk, use_cuda_graph):
    """The 30-iteration inner loop on ``st.work`` (whose last frame is about to be overwritten)."""
    st.logits.zero_()                                     # main.py:75: start from flat probabilities
    if use_cuda_graph and (st.graph is None or st.graph_topk != sample_topk):
        side = torch.cuda.Stream(device=st.work.device)
        side.wait_stream(torch.cuda.current_stream())
        keep = st.work.clone()
        with torch.cuda.stream(side):                     # warm-up outside the capture
            st.iteration(model, sample_topk)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        st.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st.graph):
            st.iteration(model, sample_topk)
        st.graph_topk = sample_topk
        st.work.copy_(keep)
        st.logits.zero_()
    for i in range(st.iterations):
        st.dyn[0:1].copy_(st.alphas[i:i + 1])
        st.dyn[1:2].fill_(float(st.calls % (1 << 24)))    # the call counter keys the Philox stream
        st.calls += 1
        if use_cuda_graph:
            st.graph.replay()
        else:
            st.iteration(model, sample_topk)
    B, _, H, W = st.work.shape
    return st.sample.view(B, H, W)


@torch.no_grad()
def sample_next_frame(model: VqVideoDiffusionModel, tokens: torch.Tensor, iterations: int = 30,
                      sample_topk: int = -1, use_cuda_graph: bool = False, seed: int = 0) -> torch.Tensor:
    """Iteratively denoise the last frame (reference ``main.py:71-111``).

    ``tokens [B,S,H,W]``; the last frame is overwritten by the first iteration (which draws from flat logits and
    re-masks all but a ``1/iterations`` fraction), the caller's tensor is left untouched.  Each iteration draws
    every position from the current (optionally top-k filtered) logits, re-masks a ``1 - (i+1)/iterations``
    fraction -- both in ONE ``wm_sample_step`` launch, no ``[P,K]`` softmax / one-hot temporaries -- and runs one
    denoiser forward.  Returns the final draw ``[B,H,W]`` (what the reference hands to ``decoder_model.decode``).
    With ``use_cuda_graph`` one iteration is captured once per (model, shape) and replayed (the loop is launch-bound
    at the reference's 8-clip evaluation batch).
    """
    st = _sampler_state(model, tokens, iterations, seed)
    st.work.copy_(tokens)
    return _denoise_last_frame(model, st, sample_topk, use_cuda_graph).clone()


@torch.no_grad()
def sample_frames(model: VqVideoDiffusionModel, tokens: torch.Tensor, num_steps: int, iterations: int = 30,
                  sample_topk: int = -1, decoder=None, use_cuda_graph: bool = True, seed: int = 0):
    """The sampling loop of ``evaluate_model`` (``main.py:62-117``): ``num_steps`` new frames, each denoised from a fully
    masked last frame in ``iterations`` mask/replace iterations, the context shifted by one frame after each
    (``:115``).  ``decoder``: a ``VqAutoEncoder`` (or anything with ``decode(tokens) -> frames``); when given, every
    sampled frame is decoded (``:113``).  Returns ``(token_frames [num_steps,B,H,W], decoded)`` with ``decoded`` the
    list of decoded frames (``None`` without a decoder).  Clips are independent: multi-GPU sampling shards the batch
    (``parallel.shard_range``) and needs no collective.
    """
    st = _sampler_state(model, tokens, iterations, seed)
    st.work.copy_(tokens)
    st.work[:, -1] = model.num_classes                    # main.py:62: destroy all information in the last frame
    frames, decoded = [], ([] if decoder is not None else None)
    for _ in range(num_steps):
        z = _denoise_last_frame(model, st, sample_topk, use_cuda_graph).clone()
        frames.append(z)
        if decoder is not None:
            decoded.append(decoder.decode(z))
        st.work[:, :-1] = st.work[:, 1:].clone()          # shift frames (the reference's overlapping in-place form raises)
    return torch.stack(frames), decoded
