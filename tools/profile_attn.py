"""Run the attention core fwd+bwd a few times on a named config (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops

cfgs = {'c3': ((32, 16, 16, 16), 8, 32, (1, 2, 2)), 'c4': ((2, 32, 32, 32), 4, 128, (2, 3, 3)),
        'c1': ((2, 8, 16, 16), 8, 32, (1, 2, 2))}
name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
(B, S, H, W), heads, d, ext = cfgs[name]
g = torch.Generator(device='cuda').manual_seed(0)
q, k, v, do = (torch.randn(B, S, H, W, heads * d, device='cuda', generator=g).bfloat16() for _ in range(4))
scale = d ** -0.5
for _ in range(iters):
    o, lse = ops.attn_forward(q, k, v, heads, ext, scale)
    ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
torch.cuda.synchronize()
# CHECK: the tensor-core kernels against the exact SIMT kernels on this very shape (a timing of wrong results is worthless)
o_s, lse_s = ops.attn_forward(q, k, v, heads, ext, scale, ops.FLAG_SIMT)
g_s = ops.attn_backward(q, k, v, o_s, lse_s, do, heads, ext, scale, ops.FLAG_SIMT)
g_t = ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
for nm, a, b_ in zip(('out', 'dq', 'dk', 'dv'), (o,) + tuple(g_t), (o_s,) + tuple(g_s)):
    d_ = (a.float() - b_.float()).abs()
    bad = int((d_ > 1e-2 * b_.float().abs().max() + 2e-2 * b_.float().abs()).sum())
    print(f'  check {nm}: max|diff|={d_.max().item():.3e} bad={bad}/{d_.numel()}' + ('  <-- MISMATCH' if bad else ''))
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(10):
    o, lse = ops.attn_forward(q, k, v, heads, ext, scale)
e1.record()
for _ in range(10):
    ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
e2.record()
torch.cuda.synchronize()
tok = B * S * H * W
print(f'{name}: fwd {e0.elapsed_time(e1) / 10:.3f} ms  bwd {e1.elapsed_time(e2) / 10:.3f} ms  '
      f'fwd+bwd tokens/s {tok / ((e0.elapsed_time(e2)) / 10 * 1e-3):.3e}')
