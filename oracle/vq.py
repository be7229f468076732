"""Oracle (CPU restatement, numpy) of ``VectorQuantizerEMA`` (``vq-video-diffusion/vq.py``).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Distances are the direct form ``sum_d (x_d - e_d)^2`` (``vq.py:30``), evaluated here
in float64 from the fp32 inputs, i.e. to ~1e-16 relative -- the exact value up to
noise far below any fp32 gap.  ``argmin`` takes the lowest index among equal
minima (``vq.py:33``; ATen's first-minimum rule).  The reference evaluates the same
sum in fp32 with an ISA-dependent accumulation order (its ``[N,L,D,K]`` temporary
is D-contiguous, so ATen reduces it with vector-lane partial sums); its argmin can
differ from the exact one only when two codes are closer than that rounding noise
(a few fp32 ulps).  ``tests/golden/make_golden.py`` records the reference's own
indices and ``tests/test_oracle_golden.py`` pins this oracle against them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

import numpy as np


@dataclass
class VQState:
    """Buffers of the reference module (``vq.py:16-20``)."""
    embedding: np.ndarray                 # [L, K, D] fp32, persistent
    cluster_size: np.ndarray              # [L, K] fp32, persistent
    activation_count: np.ndarray = field(default=None)   # [L, K] fp32, non-persistent
    accumulated_error: np.ndarray = field(default=None)  # [L, K] fp32, non-persistent

    def __post_init__(self):
        L, K, _ = self.embedding.shape
        if self.activation_count is None:
            self.activation_count = np.zeros((L, K), np.float32)
        if self.accumulated_error is None:
            self.accumulated_error = np.zeros((L, K), np.float32)


def _flat(x: np.ndarray, L: int, D: int) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32).reshape(-1, L, D)  # vq.py:27


def distances(x: np.ndarray, embedding: np.ndarray, chunk: int = 4096) -> np.ndarray:
    """``[N, L, K]`` float64 direct-form squared distances (``vq.py:30,77-82``)."""
    L, K, D = embedding.shape
    xf = _flat(x, L, D).astype(np.float64)
    e = embedding.astype(np.float64)
    out = np.empty((xf.shape[0], L, K), np.float64)
    for c0 in range(0, xf.shape[0], chunk):
        diff = xf[c0:c0 + chunk, :, None, :] - e[None]          # [n, L, K, D]
        out[c0:c0 + chunk] = np.einsum('nlkd,nlkd->nlk', diff, diff)
    return out


def encode(x: np.ndarray, embedding: np.ndarray) -> np.ndarray:
    """Nearest code per latent, lowest index on ties (``vq.py:84-87``) -> int64 ``[N, L]``."""
    return np.argmin(distances(x, embedding), axis=-1).astype(np.int64)


def decode(indices: np.ndarray, embedding: np.ndarray) -> np.ndarray:
    """Codebook gather (``vq.py:89-94``): ``indices [..., L]``-compatible -> ``[..., D]``."""
    L, K, D = embedding.shape
    idx = np.asarray(indices).reshape(-1, L)
    flat = (np.arange(L)[None, :] * K + idx).reshape(-1)
    return embedding.reshape(L * K, D)[flat].reshape(*np.shape(indices), D)


def forward(state: VQState, x: np.ndarray, training: bool, decay: float = 0.99,
            eps: float = 1e-5) -> Tuple[np.ndarray, np.ndarray, float, float, np.ndarray]:
    """``VectorQuantizerEMA.forward`` (``vq.py:25-75``), default flags
    (``simple_update=False, laplace_smoothing=True``).

    Returns ``(quantized [x.shape], encodings one-hot fp32 [N,L,K], commitment_loss,
    perplexity, indices [N,L])`` and updates ``state`` in place exactly where the
    reference has side effects: ``accumulated_error`` always (``:35-36``);
    ``activation_count``, ``cluster_size`` and ``embedding`` only when training
    (``:42-65``).  The forward *value* of the straight-through output is
    ``x + (q - x)`` evaluated in fp32 (``:70``), which differs from the gathered code
    vector ``q`` in the last ulp; it is restated with the same two fp32 roundings.
    """
    emb = state.embedding
    L, K, D = emb.shape
    xf = _flat(x, L, D)
    idx = encode(xf, emb)                                        # [N, L]
    q = decode(idx, emb).reshape(xf.shape)                       # [N, L, D]
    err = ((q.astype(np.float32) - xf) ** 2).sum(axis=2, dtype=np.float32)   # [N, L]
    for l in range(L):
        np.add.at(state.accumulated_error[l], idx[:, l], err[:, l])
    onehot = np.zeros((xf.shape[0], L, K), np.float32)
    np.put_along_axis(onehot, idx[..., None], 1.0, axis=-1)
    if training:
        counts = onehot.sum(axis=0)                              # [L, K]
        state.activation_count += counts
        dw = np.einsum('nlk,nld->lkd', onehot.astype(np.float64), xf.astype(np.float64))
        state.cluster_size[...] = state.cluster_size * decay + counts * (1 - decay)
        n = state.cluster_size.sum(axis=-1, keepdims=True)
        smoothed = (state.cluster_size + eps) / (n + K * eps) * n
        state.embedding[...] = (emb * decay + (dw / smoothed[..., None]) * (1 - decay)).astype(np.float32)
    loss = float(np.mean((q.astype(np.float64) - xf.astype(np.float64)) ** 2))
    avg = onehot.mean(axis=0, dtype=np.float64)
    perplexity = float(np.exp(-np.sum(avg * np.log(avg + 1e-10) / L)))
    ste = xf + (q.astype(np.float32) - xf)                       # two fp32 roundings, as vq.py:70
    return ste.reshape(np.shape(x)).astype(np.float32), onehot, loss, perplexity, idx


def reuse_inactive(state: VQState) -> int:
    """Move never-activated codes next to the most active ones (``vq.py:96-107``)."""
    total = 0
    for l in range(state.embedding.shape[0]):
        dead = state.activation_count[l] == 0
        nd = int(dead.sum())
        if nd:
            top = np.argsort(-state.activation_count[l], kind='stable')[:nd]
            state.embedding[l][dead] = state.embedding[l][dead] * 0.1 + state.embedding[l][top] * 0.9
            total += nd
    return total
