"""Time the attention forward (and optionally backward) kernels only -- for kernel experiments (WM_B200_LIB=variant.so)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops
cfgs = {'c3': ((32, 16, 16, 16), 8, 32, (1, 2, 2)), 'c4': ((2, 32, 32, 32), 4, 128, (2, 3, 3)),
        'c3h16': ((32, 16, 16, 16), 16, 32, (1, 2, 2)), 'c3h4': ((32, 16, 16, 16), 4, 32, (1, 2, 2))}
bwd = 'bwd' in sys.argv[1:]
names = [a for a in sys.argv[1:] if a in cfgs] or ['c3', 'c4']
for name in names:
    (B, S, H, W), heads, d, ext = cfgs[name]
    g = torch.Generator(device='cuda').manual_seed(0)
    q, k, v, do = (torch.randn(B, S, H, W, heads * d, device='cuda', generator=g).bfloat16() for _ in range(4))
    scale = d ** -0.5
    for _ in range(3):
        o, lse = ops.attn_forward(q, k, v, heads, ext, scale)
        if bwd: ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
    # CUDA graphs: the host side of a launch (ctypes, tensor-map encodes, allocations) must not bound a ~100 us kernel
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.attn_forward(q, k, v, heads, ext, scale)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf):
        for _ in range(10): ops.attn_forward(q, k, v, heads, ext, scale)
    if bwd:
        with torch.cuda.graph(gb):
            for _ in range(10): ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale)
    gf.replay(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    gf.replay()
    e[1].record()
    if bwd: gb.replay()
    e[2].record()
    torch.cuda.synchronize()
    print(f'{os.environ.get("WM_B200_LIB", "default")[-12:]} {name}: fwd {e[0].elapsed_time(e[1]) / 10:.3f} ms' + (f'  bwd {e[1].elapsed_time(e[2]) / 10:.3f} ms' if bwd else ''), flush=True)
