"""Algebraic identities the fused schedule relies on, checked on the CPU oracle (fp64 where it matters).

* ``to_v``'s bias passes through the attention core unchanged (softmax rows sum to one), so the module equals
  ``attn(q, k, x W_v^T) W_o^T + (W_o b_v + b_o)`` -- what ``Local3dAttention.forward_deferred_bias`` computes.
* The rational erf used by ``wm_bias_gelu_*`` (Abramowitz-Stegun 7.1.26) stays within 1e-6 of the exact GELU and
  of its derivative over the range activations take.
"""
import math

import numpy as np
import torch

from oracle import local3d as O


def test_value_bias_passes_through_the_attention_core():
    g = torch.Generator().manual_seed(0)
    B, S, H, W, heads, d = 1, 3, 5, 4, 2, 8
    inner = heads * d
    q, k, v = (torch.randn(B, S, H, W, inner, generator=g, dtype=torch.float64) for _ in range(3))
    b_v = torch.randn(inner, generator=g, dtype=torch.float64)
    ext = (1, 2, 1)
    base = O.attention_core(q, k, v, heads, ext, d ** -0.5)
    shifted = O.attention_core(q, k, v + b_v, heads, ext, d ** -0.5)
    out, out_shift = (t[0] if isinstance(t, tuple) else t for t in (base, shifted))
    torch.testing.assert_close(out_shift, out + b_v, rtol=1e-12, atol=1e-12)
    # ... and through the output projection: (core + b_v) W_o^T + b_o = core W_o^T + (W_o b_v + b_o)
    w_o = torch.randn(16, inner, generator=g, dtype=torch.float64)
    b_o = torch.randn(16, generator=g, dtype=torch.float64)
    lhs = out_shift @ w_o.t() + b_o
    rhs = out @ w_o.t() + torch.addmv(b_o, w_o, b_v)
    torch.testing.assert_close(lhs, rhs, rtol=1e-12, atol=1e-12)


def _device_formula(v):
    """fp32 replica of gelu_terms() in csrc/layer_ops.cu."""
    v = np.asarray(v, dtype=np.float32)
    x = np.abs(v) * np.float32(0.70710678118654752)
    t = np.float32(1.0) / (np.float32(0.3275911) * x + np.float32(1.0))
    e = np.exp(-x * x).astype(np.float32)
    p = np.float32(1.061405429) * t + np.float32(-1.453152027)
    p = p * t + np.float32(1.421413741)
    p = p * t + np.float32(-0.284496736)
    p = p * t + np.float32(0.254829592)
    erf_abs = np.float32(1.0) - p * t * e
    cdf = np.float32(0.5) * (np.float32(1.0) + np.copysign(erf_abs, v))
    pdf_v = v * np.float32(0.3989422804014327) * e
    return cdf.astype(np.float32), pdf_v.astype(np.float32)


def test_rational_erf_gelu_and_derivative_match_exact_forms():
    v = np.linspace(-10.0, 10.0, 200001)
    cdf, pdf_v = _device_formula(v)
    exact_cdf = np.array([0.5 * (1.0 + math.erf(t / math.sqrt(2.0))) for t in v])
    exact_pdf_v = v * np.exp(-0.5 * v * v) / math.sqrt(2.0 * math.pi)
    gelu_err = np.abs(v * cdf - v * exact_cdf).max()
    grad_err = np.abs((cdf + pdf_v) - (exact_cdf + exact_pdf_v)).max()
    assert gelu_err < 2e-6, gelu_err
    assert grad_err < 1e-6, grad_err
