"""Per-kernel device times of the attention core (torch profiler), config 3 / config 4 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops
from torch.profiler import profile, ProfilerActivity
cfgs = {'c3': ((32, 16, 16, 16), 8, 32, (1, 2, 2)), 'c4': ((2, 32, 32, 32), 4, 128, (2, 3, 3)),
        'c3b4': ((4, 16, 16, 16), 8, 32, (1, 2, 2)), 'c3b9': ((9, 16, 16, 16), 8, 32, (1, 2, 2)), 'c3b18': ((18, 16, 16, 16), 8, 32, (1, 2, 2))}
for name in (sys.argv[1:] or ['c3', 'c4']):
    (B, S, H, W), heads, d, ext = cfgs[name]
    g = torch.Generator(device='cuda').manual_seed(0)
    q, k, v, do = (torch.randn(B, S, H, W, heads * d, device='cuda', generator=g).bfloat16() for _ in range(4))
    scale = d ** -0.5
    for _ in range(3):
        o, lse = ops.attn_forward(q, k, v, heads, ext, scale)
        ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            o, lse = ops.attn_forward(q, k, v, heads, ext, scale)
            ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
        torch.cuda.synchronize()
    print(name)
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:5]:
        print(f'  {e.key[:70]:70s} {e.device_time_total / e.count:9.1f} us')
