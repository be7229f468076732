"""ctypes binding of ``libwm_b200.so`` (C ABI: ``include/wm_b200.h``).

The library is the product: there is no Python / PyTorch fallback.  If the shared object
is missing the import of any op raises, loudly, with the build command.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_long, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('WM_B200_LIB') or os.path.join(_HERE, '_C', 'libwm_b200.so')   # env override: tuning experiments only

DTYPE_BF16 = 0
DTYPE_FP32 = 1
FLAG_SIMT = 1

_lib = None


class WmError(RuntimeError):
    """A libwm_b200 entry point returned a negative status."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: the CUDA extension is the only implementation of this path '
            f'(no CPU/PyTorch fallback). Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `world_modelz_b200/csrc/build.sh`.')
    L = ctypes.CDLL(LIB_PATH)
    L.wm_version.restype = c_int
    L.wm_last_error.restype = c_char_p
    L.wm_l3d_attn_uses_tensor_cores.restype = c_int
    L.wm_l3d_attn_uses_tensor_cores.argtypes = [c_int] * 9
    L.wm_l3d_attn_fwd.restype = c_int
    L.wm_l3d_attn_fwd.argtypes = [c_void_p] * 5 + [c_int] * 9 + [c_float, c_int, c_int, c_void_p]
    L.wm_l3d_attn_bwd.restype = c_int
    L.wm_l3d_attn_bwd.argtypes = [c_void_p] * 10 + [c_int] * 9 + [c_float, c_int, c_int, c_void_p]
    L.wm_l3d_attn_fwd_ld.restype = c_int
    L.wm_l3d_attn_fwd_ld.argtypes = [c_void_p] * 5 + [c_long, c_long] + [c_int] * 9 + [c_float, c_int, c_int, c_void_p]
    L.wm_l3d_attn_bwd_ld.restype = c_int
    L.wm_l3d_attn_bwd_ld.argtypes = [c_void_p] * 10 + [c_long, c_long] + [c_int] * 9 + [c_float, c_int, c_int, c_void_p]
    L.wm_vq_nearest.restype = c_int
    L.wm_vq_nearest.argtypes = [c_void_p] * 5 + [c_long, c_int, c_int, c_int, c_int, c_int, c_void_p]
    L.wm_vq_distance.restype = c_int
    L.wm_vq_distance.argtypes = [c_void_p] * 3 + [c_long, c_int, c_int, c_int, c_int, c_void_p]
    L.wm_adamw_step.restype = c_int
    L.wm_adamw_step.argtypes = [c_void_p] * 5 + [c_long, c_void_p] + [c_float] * 5 + [c_int, c_void_p]
    L.wm_adamw_step_norm.restype = c_int
    L.wm_adamw_step_norm.argtypes = [c_void_p] * 5 + [c_long, c_void_p] + [c_float] * 5 + [c_int, c_void_p, c_void_p]
    L.wm_vq_stats.restype = c_int
    L.wm_vq_stats.argtypes = [c_void_p] * 6 + [c_long, c_int, c_int, c_int, c_void_p]
    L.wm_vq_onehot.restype = c_int
    L.wm_vq_onehot.argtypes = [c_void_p, c_void_p, c_long, c_int, c_void_p]
    L.wm_sample_step.restype = c_int
    L.wm_sample_step.argtypes = [c_void_p] * 3 + [c_long, c_long, c_long, c_int, c_int, c_int, c_void_p, c_uint64, c_int, c_void_p]
    L.wm_loss_hist_update.restype = c_int
    L.wm_loss_hist_update.argtypes = [c_void_p] * 4 + [c_int, c_int, c_float, c_void_p]
    L.wm_embed_pos_fwd.restype = c_int
    L.wm_embed_pos_fwd.argtypes = [c_void_p] * 6 + [c_long, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    L.wm_reduce_blocks.restype = c_int
    L.wm_reduce_blocks.argtypes = [c_long]
    L.wm_add_layernorm_fwd.restype = c_int
    L.wm_add_layernorm_fwd.argtypes = [c_void_p] * 9 + [c_long, c_int, c_float, c_int, c_void_p]
    L.wm_add_layernorm_bwd.restype = c_int
    L.wm_add_layernorm_bwd.argtypes = [c_void_p] * 11 + [c_long, c_int, c_int, c_void_p]
    L.wm_colsum.restype = c_int
    L.wm_colsum.argtypes = [c_void_p] * 3 + [c_long, c_int, c_int, c_void_p]
    L.wm_bias_gelu_fwd.restype = c_int
    L.wm_bias_gelu_fwd.argtypes = [c_void_p] * 3 + [c_long, c_int, c_int, c_void_p]
    L.wm_bias_gelu_bwd.restype = c_int
    L.wm_bias_gelu_bwd.argtypes = [c_void_p] * 6 + [c_long, c_int, c_int, c_void_p]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise WmError(f'{what} failed ({rc}): {lib().wm_last_error().decode()}')


EXPORTS = ('wm_version', 'wm_last_error', 'wm_l3d_attn_uses_tensor_cores', 'wm_l3d_attn_fwd', 'wm_l3d_attn_bwd',
           'wm_l3d_attn_fwd_ld', 'wm_l3d_attn_bwd_ld',
           'wm_vq_nearest', 'wm_vq_distance', 'wm_adamw_step', 'wm_adamw_step_norm', 'wm_vq_stats', 'wm_vq_onehot',
           'wm_sample_step', 'wm_loss_hist_update', 'wm_embed_pos_fwd', 'wm_reduce_blocks', 'wm_add_layernorm_fwd',
           'wm_add_layernorm_bwd', 'wm_colsum', 'wm_bias_gelu_fwd', 'wm_bias_gelu_bwd')
