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

    def forward(self, x):
        feats = self.transformer(x)
        return self.logit_proj(feats[:, -1])                                  # last frame only


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
    """Loss-aware diffusion-time sampler (``importance_sampling.py:5-47``): a 100-bucket
    EMA histogram of per-sample losses.  Host-side like the reference, but the loss
    feedback is consumed asynchronously (one step late) so that it never stalls the GPU.
    """

    def __init__(self, num_histogram_buckets=100, uniform_p=0.01, alpha=0.9, warmup=10, jitter=True, seed=0):
        self.n = num_histogram_buckets
        self.uniform_p, self.alpha, self.warmup, self.jitter = uniform_p, alpha, warmup, jitter
        self._weights = torch.ones(self.n)
        self._counts = torch.zeros(self.n, dtype=torch.long)
        self._gen = torch.Generator().manual_seed(seed)
        self._pending = None

    def warmed_up(self) -> bool:
        return bool((self._counts > self.warmup).all())

    def weights(self) -> torch.Tensor:
        if not self.warmed_up():
            return torch.ones(self.n)
        w = self._weights / self._weights.sum()
        return (1 - self.uniform_p) * w + self.uniform_p / self.n

    def sample(self, batch_size: int) -> torch.Tensor:
        self._drain()
        b = torch.multinomial(self.weights(), batch_size, replacement=True, generator=self._gen).float()
        if self.jitter:
            return (b + torch.rand(batch_size, generator=self._gen)) / self.n
        return b / (self.n - 1)

    def update_with_losses(self, ts: torch.Tensor, losses: torch.Tensor) -> None:
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


class DenoiserTrainer:
    """One training step of the denoiser, batch-sharded over ``world_size`` GPUs.

    Parameters live in three flat buffers: fp32 master weights, a compute-dtype shadow the
    modules read (``param.data`` are views into it) and a flat gradient buffer that one
    multi-tensor copy fills from the per-parameter gradients after backward.  A step is: zero grads, corruption,
    forward, mean CE, backward, ONE all-reduce(SUM) of the flat gradients over NCCL, ONE
    fused AdamW launch (``wm_adamw_step``) that also refreshes the shadow copy.
    """

    def __init__(self, model: VqVideoDiffusionModel, *, lr=1e-4, weight_decay=1e-7, betas=(0.9, 0.999), eps=1e-8,
                 compute_dtype=torch.bfloat16, process_group=None, use_cuda_graph=True):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError('DenoiserTrainer needs the model on a CUDA device')
        self.model = model
        self.device = params[0].device
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        n = sum(p.numel() for p in params)
        self.n_params = n
        pad = (-n) % 8
        self.master = torch.empty(n + pad, device=self.device, dtype=torch.float32)
        self.shadow = torch.zeros(n + pad, device=self.device, dtype=compute_dtype) if compute_dtype != torch.float32 else None
        self.grad = torch.zeros(n + pad, device=self.device, dtype=compute_dtype)
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.master.zero_()
        self._params = params
        self._grad_views = []
        off = 0
        for p in params:
            k = p.numel()
            self.master[off:off + k].copy_(p.detach().reshape(-1).float())
            store = self.shadow if self.shadow is not None else self.master
            store[off:off + k].copy_(p.detach().reshape(-1))
            p.data = store[off:off + k].view_as(p)
            self._grad_views.append(self.grad[off:off + k].view_as(p))
            off += k
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
    def _forward_backward(self, tokens, r):
        for p in self._params:                # autograd then ASSIGNS fresh gradients instead of launching one add per parameter
            p.grad = None
        corrupted, target = corrupt_last_frame(tokens, r, self.K)
        logits = self.model(corrupted)
        ce = F.cross_entropy(logits.reshape(-1, self.K).float(), target.reshape(-1), reduction='none')
        per_sample = ce.view(tokens.shape[0], -1).mean(dim=1)
        loss = ce.mean()
        loss.backward()
        # one multi-tensor copy gathers every parameter gradient into the flat buffer (the all-reduce / AdamW operand)
        torch._foreach_copy_(self._grad_views, [p.grad for p in self._params])
        return loss.detach(), per_sample.detach()

    def _exchange(self):
        """The path's only collective: SUM all-reduce of the flat gradient buffer (NCCL over NVLink)."""
        if self.world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg)

    def _update(self):
        ops.adamw_step(self.master, self.shadow, self.grad, self.exp_avg, self.exp_avg_sq, self.dyn, self.betas[0],
                       self.betas[1], self.eps, self.weight_decay, 1.0 / self.world)
        self.dyn[0:1] += 1.0

    def _step_body(self, tokens, r):
        out = self._forward_backward(tokens, r)
        self._exchange()
        self._update()
        return out

    def set_lr(self, lr: float) -> None:
        self.dyn[1:2].fill_(lr)

    def step(self, tokens: torch.Tensor, r: torch.Tensor):
        """``tokens [b,S,H,W]`` int64 and ``r [b]`` on the device -> (loss, per-sample loss) tensors."""
        if not self.use_cuda_graph:
            return self._step_body(tokens, r)
        if self._graph is None:
            self._capture(tokens, r)
        st_tokens, st_r, st_loss, st_ps = self._static
        st_tokens.copy_(tokens, non_blocking=True)
        st_r.copy_(r, non_blocking=True)
        graph_a, graph_b = self._graph
        graph_a.replay()
        self._exchange()          # eager, between the two graphs: the collective is never captured
        graph_b.replay()
        return st_loss, st_ps

    def _capture(self, tokens, r):
        st_tokens, st_r = tokens.clone(), r.clone().float()
        saved = (self.master.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.dyn.clone(),
                 None if self.shadow is None else self.shadow.clone())
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                 # warm-up on a side stream (allocator, cuBLAS handles, autograd)
                before = ops.launch_count()
                self._forward_backward(st_tokens, st_r)
                self._update()
                self._launches_per_step = ops.launch_count() - before
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_a):
            st_loss, st_ps = self._forward_backward(st_tokens, st_r)
        graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_b, pool=graph_a.pool()):
            self._update()
        # undo the warm-up updates so that step 1 is step 1 (capture itself executes nothing)
        self.master.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2])
        self.dyn.copy_(saved[3])
        if self.shadow is not None:
            self.shadow.copy_(saved[4])
        self._graph, self._static = (graph_a, graph_b), (st_tokens, st_r, st_loss, st_ps)

    # -- checkpoints in the reference's format (main.py:297-309, 370-410) ------------------------------
    def checkpoint(self, step: int, opt=None) -> dict:
        """The dict the reference's training loop saves: fp32 ``model_state_dict`` from the master weights and an
        ``optimizer_state_dict`` laid out like ``torch.optim.AdamW.state_dict()``."""
        from . import checkpoint as ck
        self._sync_params_from_master()
        taken = int(round(float(self.dyn[0].item()))) - 1
        shapes = ck.param_shapes(self._params)
        opt_state = ck.flat_to_adamw_state(shapes, self.exp_avg, self.exp_avg_sq, taken, lr=float(self.dyn[1].item()),
                                           betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)
        state = self.model.state_dict()
        named = dict(self.model.named_parameters())
        off = 0
        for p in self._params:                       # parameters: full-precision master copy, not the bf16 shadow
            k = p.numel()
            for name, q in named.items():
                if q is p:
                    state[name] = self.master[off:off + k].view_as(p).clone()
            off += k
        return ck.make_checkpoint(state, opt_state, step, float(self.dyn[1].item()), opt)

    def load_checkpoint(self, data: dict) -> int:
        """Resume from a dict written by the reference (or by :meth:`checkpoint`); returns the stored step."""
        from . import checkpoint as ck
        self.model.load_state_dict(data['model_state_dict'], strict=True)      # copies into the shadow views
        named = dict(self.model.named_parameters())
        sd = data['model_state_dict']
        off = 0
        for p in self._params:
            k = p.numel()
            for name, q in named.items():
                if q is p:
                    self.master[off:off + k].copy_(sd[name].reshape(-1).float())
            off += k
        if data.get('optimizer_state_dict') is not None:
            taken = ck.adamw_state_to_flat(data['optimizer_state_dict'], ck.param_shapes(self._params), self.exp_avg,
                                           self.exp_avg_sq)
            self.dyn[0:1].fill_(float(taken + 1))
        lr = data.get('lr')
        if lr:
            self.set_lr(float(lr[0] if isinstance(lr, (list, tuple)) else lr))
        return int(data.get('step', 0))

    def _sync_params_from_master(self) -> None:
        if self.shadow is not None:
            self.shadow.copy_(self.master)

    def launches_per_step(self) -> int:
        """Launches of libwm_b200 kernels per step (attention fwd/bwd, add+LayerNorm fwd/bwd, bias column sums,
        AdamW), counted by ``ops`` during the last eager warm-up step; the graphs replay exactly those."""
        return getattr(self, '_launches_per_step', 0)


def _sample_iteration(model, work, logits, alpha, K):
    """One mask/replace iteration (``main.py:79-111``): draw every position, re-mask a ``1 - alpha`` fraction,
    run the denoiser.  ``alpha`` is a 0-d device tensor so that the captured graph can be replayed."""
    B, _, H, W = work.shape
    probs = torch.softmax(logits.float(), dim=-1)
    sample = torch.multinomial(probs, 1, replacement=True).view(B, H, W)
    remask = torch.rand(B, H, W, device=work.device) > alpha
    work[:, -1] = torch.where(remask, torch.full_like(sample, K), sample)
    return sample, model(work).reshape(B * H * W, K)


@torch.no_grad()
def sample_next_frame(model: VqVideoDiffusionModel, tokens: torch.Tensor, iterations: int = 30,
                      sample_topk: int = -1, use_cuda_graph: bool = False) -> torch.Tensor:
    """Iteratively denoise the last frame (reference ``main.py:71-111``).

    ``tokens [B,S,H,W]`` with the last frame set to the mask token.  Each iteration draws
    every position from the current logits, re-masks a ``1 - (i+1)/iterations`` fraction
    and runs one denoiser forward.  Returns the final draw ``[B,H,W]`` (what the reference
    hands to ``decoder_model.decode``).  With ``use_cuda_graph`` one iteration is captured
    and replayed (the loop is launch-bound at the reference's 8-clip evaluation batch).
    """
    B, _, H, W = tokens.shape
    K = model.num_classes
    work = tokens.clone()
    logits = torch.zeros(B * H * W, K, device=tokens.device)
    sample = None
    if use_cuda_graph and sample_topk <= 0:
        # one captured iteration per (model, shape), kept on the model and replayed for every later frame
        cache = model.__dict__.setdefault('_wm_sample_graphs', {})
        key = (tuple(tokens.shape), str(tokens.device), model.training)
        if key not in cache:
            g_work, g_logits_in = tokens.clone(), torch.zeros(B * H * W, K, device=tokens.device)
            alpha = torch.zeros((), device=tokens.device)
            side = torch.cuda.Stream(device=tokens.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up outside the capture
                _sample_iteration(model, g_work.clone(), g_logits_in, alpha, K)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_sample, g_logits = _sample_iteration(model, g_work, g_logits_in, alpha, K)
            cache[key] = (graph, g_work, g_logits_in, alpha, g_sample, g_logits)
        graph, g_work, g_logits_in, alpha, g_sample, g_logits = cache[key]
        g_work.copy_(tokens)
        g_logits_in.zero_()
        for i in range(iterations):
            alpha.fill_(min(max((i + 1) / iterations, 0.0), 1.0))
            graph.replay()
            g_logits_in.copy_(g_logits)
        return g_sample.clone()
    for i in range(iterations):
        if sample_topk > 0:
            kth = torch.topk(logits, sample_topk, dim=-1).values[:, -1:]
            logits = logits.masked_fill(logits < kth, float('-inf'))
        alpha = torch.tensor(min(max((i + 1) / iterations, 0.0), 1.0), device=tokens.device)
        sample, logits = _sample_iteration(model, work, logits, alpha, K)
    return sample
