"""Layer-level kernels (fused residual-add + LayerNorm, bias-gradient column sums) against
stock PyTorch fp32 references of the same ops."""
import pytest
import torch
import torch.nn.functional as F

from world_modelz_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('dtype,rtol,atol', [(torch.float32, 1e-5, 1e-5), (torch.bfloat16, 2e-2, 2e-2)])
@pytest.mark.parametrize('rows,dim,with_delta', [(1000, 32, True), (4096, 256, True), (777, 384, False),
                                                 (130, 1024, True), (5, 8, True)])
def test_add_layernorm_forward_backward(dtype, rtol, atol, rows, dim, with_delta):
    g = torch.Generator().manual_seed(rows + dim)
    res = torch.randn(3, rows // 3 + 1, dim, generator=g)
    delta = torch.randn_like(res) if with_delta else None
    gamma = torch.rand(dim, generator=g) + 0.5
    beta = torch.randn(dim, generator=g)
    w_sum, w_y = torch.randn_like(res), torch.randn_like(res)
    # fp32 reference on the same (rounded) inputs
    cast = lambda t: None if t is None else t.detach().to(dtype).float().clone().requires_grad_(True)
    r_res, r_delta, r_gamma, r_beta = cast(res), cast(delta), cast(gamma), cast(beta)
    total = r_res if r_delta is None else (r_res + r_delta).to(dtype).float()
    if r_delta is not None:      # keep the graph through the rounding of the sum
        total = r_res + r_delta + ((r_res + r_delta).to(dtype).float() - (r_res + r_delta)).detach()
    y = F.layer_norm(total, (dim,), r_gamma, r_beta, 1e-5)
    (total * w_sum).sum().add((y * w_y).sum()).backward()
    dev = lambda t: None if t is None else t.detach().to(DEV, dtype).requires_grad_(True)
    d_res, d_delta, d_gamma, d_beta = dev(res), dev(delta), dev(gamma), dev(beta)
    o_total, o_y = ops.add_layernorm(d_res, d_delta, d_gamma, d_beta, 1e-5)
    ((o_total.float() * w_sum.to(DEV)).sum() + (o_y.float() * w_y.to(DEV)).sum()).backward()
    torch.testing.assert_close(o_total.float().cpu(), total.detach(), rtol=rtol, atol=atol)
    torch.testing.assert_close(o_y.float().cpu(), y.detach(), rtol=rtol, atol=atol)
    torch.testing.assert_close(d_res.grad.float().cpu(), r_res.grad, rtol=rtol, atol=atol * 4)
    if with_delta:
        torch.testing.assert_close(d_delta.grad.float().cpu(), r_delta.grad, rtol=rtol, atol=atol * 4)
    scale = max(1.0, (rows ** 0.5))
    torch.testing.assert_close(d_gamma.grad.float().cpu(), r_gamma.grad, rtol=rtol, atol=atol * scale)
    torch.testing.assert_close(d_beta.grad.float().cpu(), r_beta.grad, rtol=rtol, atol=atol * scale)


@pytest.mark.parametrize('dtype,rtol,atol', [(torch.float32, 1e-5, 1e-4), (torch.bfloat16, 2e-2, 5e-2)])
@pytest.mark.parametrize('rows,cin,cout', [(4096, 256, 256), (1000, 32, 48), (333, 64, 2048), (64, 16, 8)])
def test_linear_bias_gradient(dtype, rtol, atol, rows, cin, cout):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(2, rows // 2, cin, generator=g)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    dy = torch.randn(2, rows // 2, cout, generator=g)
    ref = [t.detach().to(dtype).float().clone().requires_grad_(True) for t in (x, w, b)]
    F.linear(*ref).backward(dy.to(dtype).float())
    got = [t.detach().to(DEV, dtype).requires_grad_(True) for t in (x, w, b)]
    out = ops.linear(*got)
    out.backward(dy.to(DEV, dtype))
    torch.testing.assert_close(out.float().cpu(), F.linear(*ref).detach(), rtol=rtol, atol=atol)
    for a, r in zip(got, ref):
        torch.testing.assert_close(a.grad.float().cpu(), r.grad, rtol=rtol, atol=atol * max(1.0, rows ** 0.5 / 8))


@pytest.mark.parametrize('dtype,rtol,atol', [(torch.float32, 1e-5, 1e-5), (torch.bfloat16, 2e-2, 2e-2)])
@pytest.mark.parametrize('rows,dim', [(4096, 256), (999, 48), (130, 1024)])
def test_add_layernorm_deferred_bias(dtype, rtol, atol, rows, dim):
    """sum = res + delta + delta_bias; the backward also returns the bias gradient (column sums of d sum)."""
    g = torch.Generator().manual_seed(rows * 7 + dim)
    res, delta = torch.randn(rows, dim, generator=g), torch.randn(rows, dim, generator=g)
    bias = torch.randn(dim, generator=g)
    gamma, beta = torch.rand(dim, generator=g) + 0.5, torch.randn(dim, generator=g)
    w_sum, w_y = torch.randn(rows, dim, generator=g), torch.randn(rows, dim, generator=g)
    cast = lambda t: t.detach().to(dtype).float().clone().requires_grad_(True)
    r_res, r_delta, r_bias, r_gamma, r_beta = (cast(t) for t in (res, delta, bias, gamma, beta))
    exact = r_res + r_delta + r_bias
    total = exact + (exact.to(dtype).float() - exact).detach()
    y = F.layer_norm(total, (dim,), r_gamma, r_beta, 1e-5)
    (total * w_sum).sum().add((y * w_y).sum()).backward()
    dev = lambda t: t.detach().to(DEV, dtype).requires_grad_(True)
    d_res, d_delta, d_bias, d_gamma, d_beta = (dev(t) for t in (res, delta, bias, gamma, beta))
    o_total, o_y = ops.add_layernorm(d_res, d_delta, d_gamma, d_beta, 1e-5, d_bias)
    ((o_total.float() * w_sum.to(DEV)).sum() + (o_y.float() * w_y.to(DEV)).sum()).backward()
    torch.testing.assert_close(o_total.float().cpu(), total.detach(), rtol=rtol, atol=atol)
    torch.testing.assert_close(o_y.float().cpu(), y.detach(), rtol=rtol, atol=atol)
    torch.testing.assert_close(d_res.grad.float().cpu(), r_res.grad, rtol=rtol, atol=atol * 4)
    torch.testing.assert_close(d_delta.grad.float().cpu(), r_delta.grad, rtol=rtol, atol=atol * 4)
    scale = max(1.0, rows ** 0.5)
    torch.testing.assert_close(d_bias.grad.float().cpu(), r_bias.grad, rtol=rtol, atol=atol * scale)
    torch.testing.assert_close(d_gamma.grad.float().cpu(), r_gamma.grad, rtol=rtol, atol=atol * scale)


@pytest.mark.parametrize('dtype,rtol,atol', [(torch.float32, 1e-4, 1e-4), (torch.bfloat16, 2e-2, 3e-2)])
def test_attention_module_deferred_bias_matches_forward(dtype, rtol, atol):
    """Local3dAttention.forward_deferred_bias(x, q) = (y, b) with y + b == forward(x, q), values and gradients
    (to_v's bias folded through the softmax into the output bias); same for FeedForward."""
    from world_modelz_b200.local_3d_attention import Local3dAttention, FeedForward
    torch.manual_seed(5)
    attn = Local3dAttention((1, 1, 1), 64, heads=4, dim_head=16).to(DEV, dtype)
    ff = FeedForward(64, 128).to(DEV, dtype)
    with torch.no_grad():
        attn.to_v.bias.normal_()
        attn.to_out[0].bias.normal_()
        ff.net[3].bias.normal_()
    x = torch.randn(2, 4, 6, 8, 64, device=DEV, dtype=dtype)
    q = torch.randn_like(x)
    w = torch.randn_like(x)
    for mod, args in ((attn, (x, q)), (ff, (x,))):
        names = [n for n, _ in mod.named_parameters()]
        mod.zero_grad()
        ref = mod(*args)
        (ref.float() * w.float()).sum().backward()
        ref_grads = [p.grad.float().clone() for p in mod.parameters()]
        mod.zero_grad()
        y, b = mod.forward_deferred_bias(*args)
        assert b is not None
        got = y.float() + b.float()
        (got * w.float()).sum().backward()
        torch.testing.assert_close(got, ref.float(), rtol=rtol, atol=atol)
        for n, p, r in zip(names, mod.parameters(), ref_grads):
            torch.testing.assert_close(p.grad.float(), r, rtol=rtol, atol=atol * 8, msg=lambda m, n=n: f'{n}: {m}')


@pytest.mark.parametrize('dtype,rtol,atol', [(torch.float32, 1e-5, 1e-5), (torch.bfloat16, 2e-2, 2e-2)])
@pytest.mark.parametrize('rows,cols', [(4096, 256), (1000, 48), (77, 2048), (5, 8)])
def test_bias_gelu_forward_backward(dtype, rtol, atol, rows, cols):
    """gelu(h + b) (exact erf form, FeedForward net.0/net.1) and its backward incl. the fused bias gradient."""
    g = torch.Generator().manual_seed(rows + cols)
    h, b, dy = torch.randn(rows, cols, generator=g) * 2, torch.randn(cols, generator=g), torch.randn(rows, cols, generator=g)
    r_h, r_b = (t.to(dtype).float().clone().requires_grad_(True) for t in (h, b))
    ref = F.gelu(r_h + r_b)
    ref.backward(dy.to(dtype).float())
    d_h, d_b = (t.to(DEV, dtype).requires_grad_(True) for t in (h, b))
    out = ops.bias_gelu(d_h, d_b)
    out.backward(dy.to(DEV, dtype))
    torch.testing.assert_close(out.float().cpu(), ref.detach(), rtol=rtol, atol=atol)
    torch.testing.assert_close(d_h.grad.float().cpu(), r_h.grad, rtol=rtol, atol=atol)
    torch.testing.assert_close(d_b.grad.float().cpu(), r_b.grad, rtol=rtol, atol=atol * max(1.0, rows ** 0.5))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape,dim,rows', [((2, 3, 5, 7), 64, 19), ((3, 16, 16, 16), 256, 513), ((1, 1, 2, 2), 8, 4)])
def test_embed_pos_matches_the_stock_op_chain_bit_for_bit(dtype, shape, dim, rows):
    """wm_embed_pos_fwd == embedding(tokens) + ((pos_s + pos_h) + pos_w) of local_3d_attention.py:143-157, rounded to the
    storage type in the same order; gradients of the table and of the three position tables against autograd."""
    g = torch.Generator().manual_seed(dim + rows)
    B, S, H, W = shape
    tokens = torch.randint(0, rows, shape, generator=g).to(DEV)
    mk = lambda n: torch.randn(n, dim, generator=g).to(DEV, dtype).requires_grad_(True)
    table, ps, ph, pw = mk(rows), mk(S), mk(H), mk(W)
    out = ops.embed_pos(tokens, table, ps, ph, pw)
    wgt = torch.randn(*shape, dim, generator=g).to(DEV, dtype)
    (out.float() * wgt.float()).sum().backward()
    got = [t.grad.clone() for t in (table, ps, ph, pw)]
    for t in (table, ps, ph, pw):
        t.grad = None
    pos = ps[:, None, None, :] + ph[None, :, None, :] + pw[None, None, :, :]
    ref = F.embedding(tokens, table) + pos.unsqueeze(0).expand(B, -1, -1, -1, -1)
    (ref.float() * wgt.float()).sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    tol = dict(rtol=1e-5, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-1)
    for a, t in zip(got, (table, ps, ph, pw)):
        torch.testing.assert_close(a.float(), t.grad.float(), **tol)


def test_projection_passthrough_folds_the_stream_gradient_into_the_dgrad():
    """ops.linear_passthrough(x, W) == (x @ W^T, x); its backward is one GEMM with beta = 1 (dgrad + stream gradient)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 4, 5, 64, generator=g).to(DEV, torch.bfloat16).requires_grad_(True)
    w = (torch.randn(96, 64, generator=g) * 0.1).to(DEV, torch.bfloat16).requires_grad_(True)
    gy = torch.randn(2, 3, 4, 5, 96, generator=g).to(DEV, torch.bfloat16)
    gx = torch.randn(2, 3, 4, 5, 64, generator=g).to(DEV, torch.bfloat16)
    y, xp = ops.linear_passthrough(x, w)
    ((y.float() * gy.float()).sum() + (xp.float() * gx.float()).sum()).backward()
    dx, dw = x.grad.clone(), w.grad.clone()
    x.grad = w.grad = None
    y2 = F.linear(x, w)
    ((y2.float() * gy.float()).sum() + (x.float() * gx.float()).sum()).backward()
    assert torch.equal(y, y2) and torch.equal(xp, x)
    torch.testing.assert_close(dx.float(), x.grad.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dw.float(), w.grad.float(), rtol=2e-2, atol=2e-1)
