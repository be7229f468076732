"""The C-ABI library loads and exports what include/wm_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as g
    from world_modelz_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib


def _declared():
    src = open(os.path.join(ROOT, 'include', 'wm_b200.h')).read()
    return re.findall(r'WM_API\s+[\w\s\*]+?\b(wm_\w+)\s*\(', src)


def test_header_symbols_exported():
    _lib = _ensure_built()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), f'{n} declared in wm_b200.h but not exported'
    assert set(names) == set(_lib.EXPORTS)


def test_version_and_argument_errors():
    _lib = _ensure_built()
    L = _lib.lib()
    assert L.wm_version() == 120
    # null pointers / bad shapes are rejected before any CUDA call
    rc = L.wm_l3d_attn_fwd(None, None, None, None, None, 1, 4, 4, 4, 2, 16, 1, 1, 1, 0.25, _lib.DTYPE_FP32, 0, None)
    assert rc == -1 and b'null pointer' in L.wm_last_error()
    rc = L.wm_l3d_attn_fwd(16, 16, 16, 16, 16, 1, 0, 4, 4, 2, 16, 1, 1, 1, 0.25, _lib.DTYPE_FP32, 0, None)
    assert rc == -1 and b'bad shape' in L.wm_last_error()
    rc = L.wm_l3d_attn_fwd(16, 16, 16, 16, 17, 1, 4, 4, 4, 2, 16, 1, 1, 1, 0.25, _lib.DTYPE_FP32, 0, None)
    assert rc == -1 and b'aligned' in L.wm_last_error()
    rc = L.wm_l3d_attn_fwd(16, 16, 16, 16, 16, 1, 4, 4, 4, 2, 16, 1, 1, 1, 0.25, 7, 0, None)
    assert rc == -1 and b'dtype' in L.wm_last_error()
    # token strides of the merged-projection variants: at least heads*dim_head, multiples of 16 bytes
    rc = L.wm_l3d_attn_fwd_ld(16, 16, 16, 16, 16, 0, 24, 1, 4, 4, 4, 2, 16, 1, 1, 1, 0.25, _lib.DTYPE_FP32, 0, None)
    assert rc == -1 and b'token strides' in L.wm_last_error()
    rc = L.wm_l3d_attn_bwd_ld(*([16] * 10), 0, 66, 1, 4, 4, 4, 2, 16, 1, 1, 1, 0.25, _lib.DTYPE_FP32, 0, None)
    assert rc == -1 and b'16 bytes' in L.wm_last_error()
    rc = L.wm_vq_nearest(16, 16, 16, None, None, 8, 1, 4, 6, _lib.DTYPE_FP32, 0, None)
    assert rc == -2 and b'multiple of 4' in L.wm_last_error()
    rc = L.wm_vq_nearest(16, 16, 16, None, None, 8, 1, 4, 8, _lib.DTYPE_BF16, 0, None)
    assert rc == -2
    assert L.wm_vq_nearest(None, None, None, None, None, 0, 1, 4, 8, _lib.DTYPE_FP32, 0, None) == 0   # empty input
    with pytest.raises(_lib.WmError):
        _lib.check(rc, 'wm_vq_nearest')


def test_ops_refuse_cpu_tensors():
    import torch
    from world_modelz_b200 import ops
    _ensure_built()
    x = torch.zeros(1, 2, 2, 2, 8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.local3d_attention(x, x, x, 1, (1, 1, 1))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.vq_nearest(torch.zeros(4, 1, 8), torch.zeros(1, 3, 8))
    # the layer-level wrappers refuse too (no silent stock-op path for CPU tensors)
    g = torch.ones(8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.add_layernorm(torch.zeros(3, 8), torch.zeros(3, 8), g, g)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.bias_gelu(torch.zeros(3, 8), g)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.linear(torch.zeros(3, 8), torch.zeros(8, 8), g)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.vq_onehot(torch.zeros(4, 1, dtype=torch.long), 8)


def test_new_entry_points_reject_bad_arguments():
    _lib = _ensure_built()
    L = _lib.lib()
    assert L.wm_vq_stats(None, None, None, None, None, None, 4, 1, 8, 8, None) == -1
    assert L.wm_vq_stats(None, None, None, None, None, None, 0, 1, 8, 8, None) == 0          # empty input
    assert L.wm_vq_onehot(16, 16, 4, 6, None) == -1 and b'multiple of 4' in L.wm_last_error()
    assert L.wm_sample_step(None, None, None, 0, 1, 4, 8, 0, 8, None, 0, _lib.DTYPE_FP32, None) == -1
    assert L.wm_sample_step(16, 16, None, 0, 1, 4, 8, 0, 8, 16, 0, 5, None) == -1 and b'dtype' in L.wm_last_error()
    assert L.wm_loss_hist_update(None, None, None, None, 4, 10, 0.9, None) == -1
    assert L.wm_adamw_step_norm(None, None, None, None, None, 8, None, 0.9, 0.999, 1e-8, 0.0, 1.0, _lib.DTYPE_FP32, None, None) == -1
