"""Helpers shared by the golden-fixture tests (mirrors tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def seeded(seed, *shape, kind='randn', hi=None):
    g = torch.Generator().manual_seed(seed)
    if kind == 'randn':
        return torch.randn(*shape, generator=g)
    if kind == 'rand':
        return torch.rand(*shape, generator=g)
    return torch.randint(0, hi, shape, generator=g)


def checksum(t):
    t = t.double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 977
    return np.array([t.sum().item(), (t * w).sum().item()])


def state_dict_of(fix, prefix='sd/'):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in fix.items() if k.startswith(prefix)}


def c1_state_dict():
    """Weights of the config-1 fixture (make_golden.py section 2)."""
    shapes = {'to_q.weight': (256, 256), 'to_k.weight': (256, 256), 'to_v.weight': (256, 256),
              'to_v.bias': (256,), 'to_out.0.weight': (256, 256), 'to_out.0.bias': (256,)}
    return {k: seeded(100 + i, *s) * 0.06 for i, (k, s) in enumerate(shapes.items())}
