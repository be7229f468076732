import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import ops, _lib
B, S, H, W, heads, d, ext = 32, 16, 16, 16, 8, 32, (1, 2, 2)
g = torch.Generator(device='cuda').manual_seed(0)
q, k, v = (torch.randn(B, S, H, W, heads * d, device='cuda', generator=g).bfloat16() for _ in range(3))
for _ in range(3):
    ops.attn_forward(q, k, v, heads, ext, d ** -0.5)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 32))()
rc = _lib.lib().wm_debug_read(buf)
import numpy as np
a = np.array(buf[:], dtype=np.int64).reshape(64, 32)
t0 = a[0, 0]
names = {0: 'pv:top', 1: 'pv:kv ok', 4: 'pv:p ok', 5: 'pv:issued', 2: 's:top', 3: 's:p ok', 6: 's:inputs ok', 7: 's:issued',
         14: 'cmp:HEAD top', 15: 'cmp:geometry done', 8: 'cmp:top', 9: 'cmp:o ok', 10: 'cmp:s ok', 13: 'cmp:math done', 11: 'cmp:st waited', 12: 'cmp:arrived',
         16: 'q0:arr', 17: 'q1:arr', 18: 'q2:arr', 19: 'q3:arr', 20: 'q0:top', 21: 'q1:top', 22: 'q2:top', 23: 'q3:top'}
print('rc', rc)
for t in range(0, 14 if len(sys.argv) < 2 else 0):
    ev = sorted((a[t, s] - t0, names[s]) for s in names if a[t, s] != 0)
    print(f'step {t}: ' + '  '.join(f'{n}@{c}' for c, n in ev))
for t in range(4, 13):
    ev = sorted((a[t, s] - a[4, 8], names[s]) for s in names if a[t, s] != 0)
    print(f'step {t}: ' + '  '.join(f'{n}@{c}' for c, n in ev))
d0 = np.diff(a[2:30, 8]); print('compute step period (cycles):', d0.tolist())


# ---- backward kernels
do = torch.randn(B, S, H, W, heads * d, device='cuda', generator=g).bfloat16()
o, lse = ops.attn_forward(q, k, v, heads, ext, d ** -0.5)
for _ in range(2):
    ops.attn_backward(q, k, v, o, lse, do, heads, ext, d ** -0.5)
torch.cuda.synchronize()
# ---- warp-specialised backward kernels
if hasattr(_lib.lib(), 'wm_debug_read_ws'):
    buf3 = (ctypes.c_longlong * (2 * 64 * 16))()
    _lib.lib().wm_debug_read_ws(buf3)
    b3 = np.array(buf3[:], dtype=np.int64).reshape(2, 64, 16)
    # issuing side = warp 8 (the (S,dP) issuer): iteration top, columns A drained, (S,dP) half A of the next step issued,
    # half B issued; compute side = thread 0
    nw = {0: 'iss:top', 1: 'iss:pA ok', 2: 'iss:T_A(t+1) issued', 3: 'iss:T_B(t+1) issued', 14: 'cmp:HEAD top', 15: 'cmp:geometry done', 8: 'cmp:top', 9: 'cmp:bufs free',
          10: 'cmp:T_A ready', 11: 'cmp:done', 12: 'cmp:arrived'}
    for m, name in ((0, 'dQ ws'), (1, 'dK/dV ws')):
        a = b3[m]; t0 = a[0, 0]
        print('====', name)
        for t in range(4, 13):
            ev = sorted((a[t, s] - a[4, 8], nw[s]) for s in nw if a[t, s] != 0)
            print(f'step {t}: ' + '  '.join(f'{n}@{c}' for c, n in ev))
        print('compute period:', np.diff(a[2:30, 8]).tolist())
