"""Checkpoints in the reference's on-disk format (``vq-video-diffusion/main.py:297-309,370-410``).

The reference writes ``{'step', 'lr', 'model_state_dict', 'ema_model_state_dict', 'optimizer_state_dict', 'opt'}``
with ``optimizer_state_dict`` = ``torch.optim.AdamW.state_dict()``.  ``DenoiserTrainer`` keeps its optimizer state in
three flat buffers (fp32 master weights, ``exp_avg``, ``exp_avg_sq``) and a device step counter; the functions here map
between the two, so that a run can resume from the authors' ``.pth`` files and write files their scripts can load.
Pure tensor bookkeeping, device-agnostic (covered on the CPU by ``tests/test_host.py``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


def _offsets(shapes, offsets):
    """Start of every parameter in the flat buffers: packed back to back unless the owner aligned them."""
    if offsets is not None:
        return list(offsets)
    out, off = [], 0
    for shp in shapes:
        out.append(off)
        off += int(torch.Size(shp).numel())
    return out


def flat_to_adamw_state(shapes: Sequence[torch.Size], exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, *,
                        lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                        offsets: Optional[Sequence[int]] = None) -> Dict:
    """Flat moment buffers -> ``torch.optim.AdamW.state_dict()`` layout (parameter order = ``shapes`` order;
    ``offsets``: where each parameter starts in the flat buffers, default packed)."""
    state = {}
    for i, (shp, off) in enumerate(zip(shapes, _offsets(shapes, offsets))):
        k = int(torch.Size(shp).numel())
        if step > 0:
            state[i] = {'step': torch.tensor(float(step)),
                        'exp_avg': exp_avg[off:off + k].detach().float().reshape(shp).cpu().clone(),
                        'exp_avg_sq': exp_avg_sq[off:off + k].detach().float().reshape(shp).cpu().clone()}
    group = {'lr': lr, 'betas': tuple(betas), 'eps': eps, 'weight_decay': weight_decay, 'amsgrad': False,
             'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
             'params': list(range(len(shapes)))}
    return {'state': state, 'param_groups': [group]}


def adamw_state_to_flat(opt_state: Dict, shapes: Sequence[torch.Size], exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                        offsets: Optional[Sequence[int]] = None) -> int:
    """``AdamW.state_dict()`` -> the flat moment buffers (in place).  Returns the number of steps taken (0 when the
    optimizer had not stepped).  Raises ``ValueError`` on a parameter count / shape mismatch."""
    params = opt_state['param_groups'][0]['params'] if len(opt_state['param_groups']) == 1 else \
        [p for g in opt_state['param_groups'] for p in g['params']]
    if len(params) != len(shapes):
        raise ValueError(f'optimizer state has {len(params)} parameters, the model has {len(shapes)}')
    steps = 0
    for idx, shp, off in zip(params, shapes, _offsets(shapes, offsets)):
        k = int(torch.Size(shp).numel())
        st = opt_state['state'].get(idx)
        if st is None:
            exp_avg[off:off + k].zero_()
            exp_avg_sq[off:off + k].zero_()
        else:
            if tuple(st['exp_avg'].shape) != tuple(shp):
                raise ValueError(f'parameter {idx}: optimizer state shape {tuple(st["exp_avg"].shape)} != {tuple(shp)}')
            exp_avg[off:off + k].copy_(st['exp_avg'].reshape(-1))
            exp_avg_sq[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
            steps = max(steps, int(float(st['step'])))
    return steps


def make_checkpoint(model_state: Dict[str, torch.Tensor], optimizer_state: Dict, step: int, lr: float,
                    opt=None, ema_model_state: Optional[Dict[str, torch.Tensor]] = None) -> Dict:
    """The dict ``main.py:300-307`` hands to ``torch.save``."""
    return {'step': step, 'lr': [lr], 'model_state_dict': {k: v.detach().float().cpu() for k, v in model_state.items()},
            'ema_model_state_dict': ema_model_state, 'optimizer_state_dict': optimizer_state, 'opt': opt}


def param_shapes(params: Sequence[torch.Tensor]) -> List[torch.Size]:
    return [p.shape for p in params]
