"""Parity of the CUDA attention kernels (through the C ABI) against the oracle.

Bars (BASELINE.json north_star): fp32 path rtol 1e-5; bf16 path rtol 2e-2.  Tolerances
are stated as |a-b| <= atol + rtol*|b| with atol a small fraction of the tensor's scale
(the outputs are sums of O(100) products, so individual elements cross zero).
"""
import numpy as np
import pytest
import torch

import world_modelz_b200 as wm
from world_modelz_b200 import ops
from oracle import local3d as O
from tests._golden import load, seeded, checksum, state_dict_of, c1_state_dict

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _close(a, b, rtol, atol_frac, what=''):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    atol = atol_frac * b.abs().max().item() + 1e-6      # 1e-6: analytically-zero gradients come out as fp32 noise
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), (f'{what}: {int(bad.sum())}/{bad.numel()} elements off; max|diff|='
                           f'{(a - b).abs().max().item():.3e} scale={b.abs().max().item():.3e}')


def _core_case(shape, heads, ext, dtype, seed=0, flags=0):
    g = torch.Generator().manual_seed(seed)
    q, k, v, do = (torch.randn(shape, generator=g) for _ in range(4))
    if dtype == torch.bfloat16:      # compare against the oracle on the SAME (rounded) inputs
        q, k, v, do = (t.bfloat16().float() for t in (q, k, v, do))
    leaves = [t.clone().requires_grad_(True) for t in (q, k, v)]
    ref, ref_lse = O.attention_core(*leaves, heads, ext, want_lse=True)
    ref.backward(do)
    dl = [t.to(DEV, dtype).requires_grad_(True) for t in (q, k, v)]
    out = ops.local3d_attention(*dl, heads, ext, flags=flags)
    out.backward(do.to(DEV, dtype))
    _, lse = ops.attn_forward(*[t.detach() for t in dl], heads, ext, (shape[-1] // heads) ** -0.5, flags)
    torch.cuda.synchronize()
    return out, lse, [t.grad for t in dl], ref, ref_lse, [t.grad for t in leaves]


FP32_CASES = [
    ((2, 4, 6, 5, 32), 2, (1, 2, 2)),
    ((1, 3, 3, 3, 16), 1, (2, 2, 2)),      # window larger than the grid on every axis
    ((1, 1, 7, 9, 24), 3, (1, 1, 1)),      # single frame
    ((2, 5, 4, 4, 8), 2, (0, 0, 0)),       # window of one key: out == v
    ((1, 6, 5, 7, 128), 1, (2, 3, 3)),     # config-4 head width, ragged grid
    ((1, 2, 16, 16, 64), 2, (3, 1, 1)),    # the (7,3,3) window of the published results
]


@pytest.mark.parametrize('shape,heads,ext', FP32_CASES)
def test_fp32_core_forward_backward(shape, heads, ext):
    out, lse, grads, ref, ref_lse, ref_grads = _core_case(shape, heads, ext, torch.float32)
    _close(out, ref, 1e-5, 2e-6, 'out')
    _close(lse, ref_lse, 1e-5, 2e-6, 'lse')
    for name, g, rg in zip('qkv', grads, ref_grads):
        _close(g, rg, 1e-5, 5e-6, 'd' + name)


@pytest.mark.parametrize('flags', [0, ops.FLAG_SIMT])
@pytest.mark.parametrize('shape,heads,ext', [
    ((2, 4, 16, 16, 64), 2, (1, 2, 2)),     # config-1/3 head width (32), exact tiling
    ((1, 6, 10, 10, 256), 2, (2, 3, 3)),    # config-4 head width (128), ragged tiles
    ((1, 3, 9, 7, 64), 1, (1, 1, 2)),       # dim_head 64, grid smaller than a tile
    ((1, 8, 8, 8, 32), 1, (3, 1, 1)),
    ((1, 5, 12, 20, 96), 3, (0, 2, 1)),
])
def test_bf16_core_forward_backward(shape, heads, ext, flags):
    out, lse, grads, ref, ref_lse, ref_grads = _core_case(shape, heads, ext, torch.bfloat16, flags=flags)
    _close(out, ref, 2e-2, 4e-3, 'out')
    _close(lse, ref_lse, 2e-2, 2e-3, 'lse')
    for name, g, rg in zip('qkv', grads, ref_grads):
        _close(g, rg, 2e-2, 8e-3, 'd' + name)


def test_bf16_tensor_core_path_is_selected_for_the_named_configs():
    assert ops.uses_tensor_cores(8, 16, 16, 8, 32, (1, 2, 2))       # config 1
    assert ops.uses_tensor_cores(16, 16, 16, 8, 32, (1, 2, 2))      # config 3
    assert ops.uses_tensor_cores(32, 32, 32, 4, 128, (2, 3, 3))     # config 4
    assert not ops.uses_tensor_cores(8, 16, 16, 8, 32, (1, 2, 2), torch.float32)


@pytest.mark.parametrize('name', ['attn_small.npz', 'attn_noproj.npz'])
def test_module_against_reference_fixture(name):
    """The drop-in module, loaded with the reference's state_dict, reproduces the
    reference's own outputs and gradients (fixture written by the unmodified reference)."""
    f = load(name)
    m = wm.Local3dAttention(tuple(int(e) for e in f['ext']), f['x'].shape[-1], heads=int(f['heads']),
                            dim_head=int(f['dim_head'])).to(DEV)
    m.load_state_dict(state_dict_of(f))
    x = torch.from_numpy(f['x']).to(DEV).requires_grad_(True)
    q = torch.from_numpy(f['q']).to(DEV).requires_grad_(True)
    out = m(x, q)
    out.backward(torch.from_numpy(f['dout']).to(DEV))
    _close(out, torch.from_numpy(f['out']), 1e-5, 2e-6, 'out')
    _close(x.grad, torch.from_numpy(f['dx']), 1e-5, 5e-6, 'dx')
    _close(q.grad, torch.from_numpy(f['dq']), 1e-5, 5e-6, 'dq')
    for k, p in m.named_parameters():
        if 'grad/' + k in f:
            _close(p.grad, torch.from_numpy(f['grad/' + k]), 1e-4, 1e-5, k)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_config1_module_against_reference_fixture(dtype):
    """BASELINE config 1: B2, 8x16x16 tokens, dim 256, 8 heads x 32, window 3x5x5."""
    f = load('attn_c1.npz')
    x = seeded(7, 2, 8, 16, 16, 256)
    dout = seeded(8, 2, 8, 16, 16, 256)
    if not np.allclose(checksum(x), f['x_sum'], rtol=1e-12):
        pytest.skip('seeded inputs differ from the fixture')
    m = wm.Local3dAttention((1, 2, 2), 256, heads=8, dim_head=32).to(DEV)
    m.load_state_dict(c1_state_dict())
    m = m.to(dtype)
    xd = x.to(DEV, dtype).requires_grad_(True)
    out = m(xd, xd)
    out.backward(dout.to(DEV, dtype))
    tok = slice(5, None, 16)
    rtol, af = (1e-5, 5e-6) if dtype == torch.float32 else (2e-2, 1e-2)
    _close(out.reshape(-1, 256)[tok], torch.from_numpy(f['out']), rtol, af, 'out')
    _close(xd.grad.reshape(-1, 256)[tok], torch.from_numpy(f['dx']), 10 * rtol if dtype == torch.float32 else rtol,
           af * 2, 'dx')
    if dtype == torch.float32:
        for k, p in m.named_parameters():
            _close(p.grad.flatten()[::97], torch.from_numpy(f['grad/' + k]), 1e-4, 1e-5, k)


def test_config4_core_reduced_grid_fixture():
    f = load('core_c4_reduced.npz')
    shape = tuple(int(s) for s in f['shape'])
    q, k, v, do = seeded(11, *shape) * 0.5, seeded(12, *shape) * 0.5, seeded(13, *shape), seeded(14, *shape)
    if not np.allclose(checksum(q), f['q_sum'], rtol=1e-12):
        pytest.skip('seeded inputs differ from the fixture')
    for dtype, rtol, af in ((torch.float32, 1e-5, 5e-6), (torch.bfloat16, 2e-2, 1e-2)):
        dl = [t.to(DEV, dtype).requires_grad_(True) for t in (q, k, v)]
        out = ops.local3d_attention(*dl, int(f['heads']), tuple(f['ext']))
        out.backward(do.to(DEV, dtype))
        _close(out[0, ::2, ::3, ::3], torch.from_numpy(f['out']), rtol, af, f'out {dtype}')
        for name, t in zip(('dq', 'dk', 'dv'), dl):
            _close(t.grad[0, ::2, ::3, ::3], torch.from_numpy(f[name]), rtol, 2 * af, f'{name} {dtype}')


def test_full_size_config4_against_oracle_slabs():
    """BASELINE config 4 at FULL size -- 32x32x32 tokens, dim 512 (4 heads x 128), window 5x7x7 -- forward and all
    three gradients against the CPU oracle on query slabs (planes around s = 0, 15, 31: both grid borders and the
    interior).  ``oracle.attention_core_slab`` evaluates only the slab's queries; dK / dV are compared on the slab's
    middle plane, whose observers all lie inside the slab."""
    shape, heads, ext = (1, 32, 32, 32, 512), 4, (2, 3, 3)
    S, H, W = shape[1:4]
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(shape, generator=g) * 0.3).bfloat16()
    k = (torch.randn(shape, generator=g) * 0.3).bfloat16()
    v = torch.randn(shape, generator=g).bfloat16()
    do = torch.randn(shape, generator=g).bfloat16()
    qd, kd, vd, dod = (t.to(DEV) for t in (q, k, v, do))
    scale = 128 ** -0.5
    o_tc, lse_tc = ops.attn_forward(qd, kd, vd, heads, ext, scale)
    dq_tc, dk_tc, dv_tc = ops.attn_backward(qd, kd, vd, o_tc, lse_tc, dod, heads, ext, scale)
    torch.cuda.synchronize()
    plane = H * W
    for centre in (0, 15, 31):
        s_lo, s_hi = max(0, centre - ext[0]), min(S - 1, centre + ext[0])           # query planes of the slab
        k_lo, k_hi = max(0, s_lo - ext[0]), min(S - 1, s_hi + ext[0])               # planes their windows touch
        sub = [t[:, k_lo:k_hi + 1].float() for t in (q, k, v, do)]
        ids = torch.arange((s_lo - k_lo) * plane, (s_hi - k_lo + 1) * plane)
        out, dq, dk, dv = O.attention_core_slab(*sub, heads, ext, ids, scale=scale, chunk=256)
        C = shape[-1]
        _close(o_tc[0, s_lo:s_hi + 1].reshape(-1, C), out[0], 2e-2, 4e-3, f'out, planes {s_lo}..{s_hi}')
        _close(dq_tc[0, s_lo:s_hi + 1].reshape(-1, C), dq[0], 2e-2, 1e-2, f'dq, planes {s_lo}..{s_hi}')
        _close(dk_tc[0, centre], dk[0, centre - k_lo], 2e-2, 1e-2, f'dk, plane {centre}')
        _close(dv_tc[0, centre], dv[0, centre - k_lo], 2e-2, 1e-2, f'dv, plane {centre}')


def test_full_size_properties_config4():
    """Size-independent properties at the full config-4 size (the oracle comparison is the test above): (i) the
    tensor-core and SIMT kernels agree everywhere, (ii) constant V rows give constant outputs (softmax weights sum
    to 1), (iii) gradients of sum(out) w.r.t. q vanish and dv sums to the number of queries."""
    shape, heads, ext = (1, 32, 32, 32, 512), 4, (2, 3, 3)
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(shape, generator=g) * 0.3).to(DEV, torch.bfloat16)
    k = (torch.randn(shape, generator=g) * 0.3).to(DEV, torch.bfloat16)
    v = torch.randn(shape, generator=g).to(DEV, torch.bfloat16)
    scale = 128 ** -0.5
    o_tc, lse_tc = ops.attn_forward(q, k, v, heads, ext, scale)
    o_si, lse_si = ops.attn_forward(q, k, v, heads, ext, scale, ops.FLAG_SIMT)
    _close(o_tc, o_si, 2e-2, 4e-3, 'tc vs simt out')
    _close(lse_tc, lse_si, 1e-3, 1e-4, 'tc vs simt lse')
    do = torch.randn(shape, generator=g).to(DEV, torch.bfloat16)
    g_tc = ops.attn_backward(q, k, v, o_tc, lse_tc, do, heads, ext, scale)
    g_si = ops.attn_backward(q, k, v, o_si, lse_si, do, heads, ext, scale, ops.FLAG_SIMT)
    for name, a, b in zip(('dq', 'dk', 'dv'), g_tc, g_si):
        _close(a, b, 2e-2, 1e-2, 'tc vs simt ' + name)
    ones = torch.ones_like(v)
    o1, l1 = ops.attn_forward(q, k, ones, heads, ext, scale)
    assert (o1.float() - 1).abs().max().item() < 1e-2
    dq, dk, dv = ops.attn_backward(q, k, ones, o1, l1, ones, heads, ext, scale)
    assert dq.float().abs().max().item() < 2e-2 * max(1.0, k.float().abs().max().item())
    n_queries = shape[1] * shape[2] * shape[3]
    assert abs(dv.float()[..., 0].sum().item() / n_queries - 1) < 1e-2


def test_rows_outside_the_max_free_range_are_recomputed_exactly():
    """The tensor-core forward computes 2^(s*scale*log2e) against a FIXED exponent; rows whose sum leaves
    [2^-100, 2^100] are marked (LSE = NaN) and recomputed by the exact fix-up kernel.  Mix ordinary rows with rows of
    +-150-nat logits in one launch: no NaN may survive and everything matches the exact SIMT kernels."""
    shape, heads, ext = (1, 4, 16, 16, 64), 2, (1, 2, 2)
    g = torch.Generator().manual_seed(9)
    q, k, v = (torch.randn(shape, generator=g).bfloat16() for _ in range(3))
    qf = q.float()
    qf[0, 1, :8] *= 60.0                  # huge positive and negative logits
    qf[0, 2, 8:, :4] *= -45.0
    q = qf.bfloat16()
    qd, kd, vd = (t.to(DEV) for t in (q, k, v))
    o_si, l_si = ops.attn_forward(qd, kd, vd, heads, ext, 32 ** -0.5, ops.FLAG_SIMT)
    o_tc, l_tc = ops.attn_forward(qd, kd, vd, heads, ext, 32 ** -0.5, 0)
    assert not torch.isnan(l_tc).any() and not torch.isnan(o_tc.float()).any()
    assert l_si.abs().max().item() > 100
    _close(o_tc, o_si, 2e-2, 4e-3, 'out')
    _close(l_tc, l_si, 1e-3, 1e-4, 'lse')


def test_unsupported_dtype_and_shape_raise():
    x = torch.zeros(1, 2, 2, 2, 16, device=DEV, dtype=torch.float16)
    with pytest.raises(TypeError):
        ops.local3d_attention(x, x, x, 1, (1, 1, 1))
    y = torch.zeros(1, 2, 2, 2, 6, device=DEV)
    with pytest.raises(RuntimeError, match='dim_head'):
        ops.local3d_attention(y, y, y, 1, (1, 1, 1))


@pytest.mark.parametrize('gain', [12.0, 40.0])
def test_bf16_extreme_logits_recentre_the_softmax_reference(gain):
    """Logits of +-50 .. +-180 nats with later planes scoring far lower than the first ones: exercises the
    re-centring (two-pass, O / l rescale in TMEM) branch of the tensor-core softmax against the exact kernels."""
    shape, heads, ext = (1, 4, 16, 16, 64), 2, (1, 2, 2)
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(shape, generator=g).bfloat16() for _ in range(3))
    q = (q.float() * gain).bfloat16()
    k[:, 2:] = (k[:, 2:].float() * 0.05).bfloat16()
    qd, kd, vd = (t.to(DEV) for t in (q, k, v))
    o_si, l_si = ops.attn_forward(qd, kd, vd, heads, ext, 32 ** -0.5, ops.FLAG_SIMT)
    o_tc, l_tc = ops.attn_forward(qd, kd, vd, heads, ext, 32 ** -0.5, 0)
    assert l_si.abs().max().item() > 40
    _close(o_tc, o_si, 2e-2, 4e-3, 'out')
    _close(l_tc, l_si, 1e-3, 1e-4, 'lse')
    ref = O.attention_core(q.float(), k.float(), v.float(), heads, ext)
    _close(o_tc, ref, 2e-2, 6e-3, 'out vs oracle')


@pytest.mark.parametrize('dtype,flags', [(torch.bfloat16, 0), (torch.bfloat16, ops.FLAG_SIMT), (torch.float32, 0)])
@pytest.mark.parametrize('shape,heads,ext', [
    ((2, 4, 16, 16, 64), 2, (1, 2, 2)),     # config-3 head width
    ((1, 5, 10, 9, 256), 2, (2, 3, 3)),     # config-4 head width, ragged tiles
])
def test_merged_kv_projection_operands_match_separate_tensors(shape, heads, ext, dtype, flags):
    """wm_l3d_attn_*_ld on the two channel halves of ONE [.., 2*inner] buffer (merged to_k / to_v GEMM) gives the same
    bits as the contiguous entry points on copies of the halves, forward and backward (dK | dV side by side)."""
    g = torch.Generator().manual_seed(5)
    C = shape[-1]
    q = torch.randn(shape, generator=g).to(DEV, dtype).requires_grad_(True)
    kv = torch.randn(*shape[:-1], 2 * C, generator=g).to(DEV, dtype).requires_grad_(True)
    do = torch.randn(shape, generator=g).to(DEV, dtype)
    out = ops.local3d_attention_kv(q, kv, heads, ext, flags=flags)
    out.backward(do)
    q2 = q.detach().clone().requires_grad_(True)
    k2 = kv.detach()[..., :C].contiguous().requires_grad_(True)
    v2 = kv.detach()[..., C:].contiguous().requires_grad_(True)
    ref = ops.local3d_attention(q2, k2, v2, heads, ext, flags=flags)
    ref.backward(do)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert torch.equal(q.grad, q2.grad)
    assert torch.equal(kv.grad[..., :C], k2.grad) and torch.equal(kv.grad[..., C:], v2.grad)


@pytest.mark.parametrize('shape,heads,ext', [
    ((5, 16, 16, 16, 256), 8, (1, 2, 2)),      # config-3 widths: 160 work items > 148 SMs, some CTAs walk two items
    ((32, 16, 16, 16, 256), 8, (1, 2, 2)),     # the benchmarked shape: 6-7 items per persistent CTA, s-border and interior mixed
    ((3, 12, 40, 24, 64), 2, (1, 2, 2)),       # many brick positions (5 x 3), ragged in s, one head per item
])
def test_persistent_grid_matches_the_exact_kernels(shape, heads, ext):
    """Shapes with more work items than SMs run the tensor-core kernels on their persistent grid (a CTA walks several
    (batch, s position, head group) items of one brick position): forward and all gradients against the exact SIMT
    kernels on the same inputs, everywhere."""
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(shape, generator=g) * 0.5).to(DEV, torch.bfloat16)
    k = (torch.randn(shape, generator=g) * 0.5).to(DEV, torch.bfloat16)
    v = torch.randn(shape, generator=g).to(DEV, torch.bfloat16)
    do = torch.randn(shape, generator=g).to(DEV, torch.bfloat16)
    scale = (shape[-1] // heads) ** -0.5
    assert ops.uses_tensor_cores(*shape[1:4], heads, shape[-1] // heads, ext)
    o_tc, lse_tc = ops.attn_forward(q, k, v, heads, ext, scale)
    o_si, lse_si = ops.attn_forward(q, k, v, heads, ext, scale, ops.FLAG_SIMT)
    _close(o_tc, o_si, 2e-2, 4e-3, 'out')
    _close(lse_tc, lse_si, 1e-3, 1e-4, 'lse')
    g_tc = ops.attn_backward(q, k, v, o_tc, lse_tc, do, heads, ext, scale)
    g_si = ops.attn_backward(q, k, v, o_si, lse_si, do, heads, ext, scale, ops.FLAG_SIMT)
    for name, a, b in zip(('dq', 'dk', 'dv'), g_tc, g_si):
        _close(a, b, 2e-2, 1e-2, name)
