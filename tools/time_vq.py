"""Time and check the VQ nearest kernel (tensor-core path vs the exact SIMT path) at the benchmark size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn(n, 1, 64, device='cuda', generator=g)
cb = torch.randn(1, 512, 64, device='cuda', generator=g)
i_tc, q_tc, e_tc = ops.vq_nearest(x, cb)
i_si, q_si, e_si = ops.vq_nearest(x, cb, flags=ops.FLAG_SIMT)
torch.cuda.synchronize()
print('index mismatches vs exact kernel:', int((i_tc != i_si).sum()), 'of', n, ' q equal:', bool(torch.equal(q_tc, q_si)),
      ' err max diff:', float((e_tc - e_si).abs().max()))
for name, fl in (('tensor-core', 0), ('simt', ops.FLAG_SIMT)):
    for _ in range(2): ops.vq_nearest(x, cb, flags=fl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.vq_nearest(x, cb, flags=fl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'{name}: {ms:.3f} ms  {n / ms / 1e6:.2f} G latents/s  {n * 524 / ms / 1e6:.0f} GB/s algorithmic')
