"""Time the layer-level kernels (fused add+LayerNorm fwd/bwd, bias column sums) at the config-3 shape, flushing L2
between launches, and report achieved HBM GB/s on their algorithmic bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from world_modelz_b200 import _lib
from world_modelz_b200.ops import check, _stream

rows, dim = 32 * 16 * 16 * 16, 256
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
res, delta, dy, dres = rnd(rows, dim), rnd(rows, dim), rnd(rows, dim), rnd(rows, dim)
gamma, beta = rnd(dim), rnd(dim)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
total, y = torch.empty_like(res), torch.empty_like(res)
mean, rstd = (torch.empty(rows, device=dev, dtype=torch.float32) for _ in range(2))
dx, dgamma, dbeta = torch.empty_like(res), torch.empty_like(gamma), torch.empty_like(gamma)
ws = torch.empty(L.wm_reduce_blocks(rows) * 3 * dim, device=dev, dtype=torch.float32)
db = torch.empty(dim, device=dev, dtype=torch.bfloat16)

def fwd():
    check(L.wm_add_layernorm_fwd(res.data_ptr(), delta.data_ptr(), beta.data_ptr(), gamma.data_ptr(), beta.data_ptr(), total.data_ptr(), y.data_ptr(),
                                 mean.data_ptr(), rstd.data_ptr(), rows, dim, 1e-5, 0, _stream()), 'fwd')
def bwd():
    check(L.wm_add_layernorm_bwd(dy.data_ptr(), dres.data_ptr(), total.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                 dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), db.data_ptr(), ws.data_ptr(), rows, dim, 0, _stream()), 'bwd')
def colsum():
    check(L.wm_colsum(dy.data_ptr(), db.data_ptr(), ws.data_ptr(), rows, dim, 0, _stream()), 'colsum')

hid, bias_h = rnd(rows, dim), rnd(dim)
gy, gdh, gdb = torch.empty_like(hid), torch.empty_like(hid), torch.empty_like(bias_h)
def gelu_fwd():
    check(L.wm_bias_gelu_fwd(hid.data_ptr(), bias_h.data_ptr(), gy.data_ptr(), rows, dim, 0, _stream()), 'gelu fwd')
def gelu_bwd():
    check(L.wm_bias_gelu_bwd(dy.data_ptr(), hid.data_ptr(), bias_h.data_ptr(), gdh.data_ptr(), gdb.data_ptr(), ws.data_ptr(), rows, dim, 0,
                             _stream()), 'gelu bwd')

def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n

fwd()
t = rows * dim * 2
for name, fn, nbytes in (('add_layernorm_fwd', fwd, 4 * t), ('add_layernorm_bwd (+reduce)', bwd, 4 * t), ('colsum (+reduce)', colsum, t),
                        ('bias_gelu_fwd', gelu_fwd, 2 * t), ('bias_gelu_bwd (+reduce)', gelu_bwd, 3 * t)):
    ms = timeit(fn)
    print(f'{name:30s} {ms * 1e3:7.1f} us   {nbytes / ms / 1e6:7.0f} GB/s algorithmic')
