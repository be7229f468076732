"""Timeline of CTA 0 of vq_nearest_tc_kernel (library built with -DWM_VQ_EXP=32):
  WM_B200_LIB=world_modelz_b200/_C/libwm_vqtl.so python tools/vq_timeline.py
Events per tile (clock64 of the SM, relative to the first event of tile 20):
  writer warp 8:  0 convert: start   1 fp32 rows landed   2 operands written (xready)   3 buffer free (MMAs of tile-2 retired), copy issued
                  4 output: start    5 scan results there 6 output written
  scanner half A: 7 scores of its first quarter there    8 scan done    11 results posted;   half B: 12, 13
  MMA issuer:     9 operands of the tile written (xready)   10 the tile's four chains issued
  writer warp 8:  15 output pass 1 done (winner per row known, idx written)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops, _lib
n = 1 << 20
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn(n, 1, 64, device='cuda', generator=g)
cb = torch.randn(1, 512, 64, device='cuda', generator=g)
for _ in range(3):
    ops.vq_nearest(x, cb)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (64 * 24))()
L = _lib.lib()
L.wm_vq_debug_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
assert L.wm_vq_debug_read(buf) == 0
names = ['cv0', 'land', 'xrdy', 'xfree', 'out0', 'res', 'outE', 'A.full', 'A.scanE', 'iss.xrdy', 'iss.done', 'A.post', 'B.full', 'B.scanE', 'out.r0', 'out.p1', 'out.r7']
t0 = min(buf[20 * 24 + e] for e in range(17) if buf[20 * 24 + e])
print('tile ' + ' '.join(f'{s:>8}' for s in names))
for j in range(20, 30):
    print(f'{j:4d} ' + ' '.join(f'{(buf[j * 24 + e] - t0) if buf[j * 24 + e] else -1:8d}' for e in range(17)))
