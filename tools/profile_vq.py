import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(1 << 20, 1, 64, device='cuda', generator=g)
cb = torch.randn(1, 512, 64, device='cuda', generator=g)
for _ in range(3):
    ops.vq_nearest(x, cb)
torch.cuda.synchronize()
