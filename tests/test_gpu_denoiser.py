"""Denoiser (transformer + head), training step and sampler on the GPU against the oracle."""
import numpy as np
import pytest
import torch

import world_modelz_b200 as wm
from oracle import local3d as O
from tests._golden import load, state_dict_of

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _small(f):
    c = [int(v) for v in f['cfg']]
    kw = dict(data_shape=tuple(c[0:3]), dim=c[3], num_classes=c[4], extents=tuple(c[5:8]), depth=c[8], heads=c[9],
              dim_head=c[10], mlp_dim=c[11])
    return kw, O.DenoiserConfig(**kw)


@pytest.mark.parametrize('prune', [False, True])
def test_denoiser_against_reference_fixture_fp32(prune):
    """prune=True: the opt-in receptive-field cone (4 frames, depth 2, extents (1, ., .): layers see frames 1.., 2..)
    must give the reference's logits, loss and parameter gradients -- including the exact zeros of the position rows
    and embedding rows only dead frames touch."""
    f = load('denoiser_small.npz')
    kw, _ = _small(f)
    m = wm.VqVideoDiffusionModel(**kw).to(DEV)
    m.load_state_dict(state_dict_of(f))
    m.prune_receptive_field = prune
    assert m.transformer.last_frame_cone(4) == [1, 2, 3]
    tokens = torch.from_numpy(f['tokens']).to(DEV)
    target = torch.from_numpy(f['target']).to(DEV)
    feats = m.transformer(tokens)
    np.testing.assert_allclose(feats.detach().cpu().numpy(), f['feats'], rtol=1e-4, atol=2e-5)
    logits = m(tokens)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), f['logits'], rtol=1e-4, atol=2e-5)
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, 17), target.reshape(-1))
    assert abs(loss.item() - float(f['loss'])) < 1e-5
    loss.backward()
    for k, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), f['grad/' + k], rtol=2e-4, atol=5e-6, err_msg=k)


@pytest.mark.parametrize('graph', [False, True])
def test_training_steps_match_oracle_fp32(graph):
    """Three AdamW steps in exact (fp32) mode reproduce the oracle's train_step: same
    losses and same updated weights.  Corruption is RNG-dependent, so r=0 (no corruption)."""
    cfg = O.DenoiserConfig(data_shape=(3, 8, 8), dim=32, num_classes=20, extents=(1, 2, 1), depth=2, heads=2,
                           dim_head=16, mlp_dim=40)
    p = O.init_denoiser_params(cfg, seed=3)
    m = wm.VqVideoDiffusionModel(data_shape=cfg.data_shape, dim=cfg.dim, num_classes=cfg.num_classes,
                                 extents=cfg.extents, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head,
                                 mlp_dim=cfg.mlp_dim).to(DEV)
    m.load_state_dict(p)
    tr = wm.DenoiserTrainer(m, lr=1e-2, weight_decay=1e-2, compute_dtype=torch.float32, use_cuda_graph=graph)
    ref_p = {k: v.clone() for k, v in p.items()}
    state = {}
    g = torch.Generator().manual_seed(0)
    for step in range(1, 4):
        tokens = torch.randint(0, 20, (4, 3, 8, 8), generator=g)
        ref_loss = O.train_step(ref_p, state, step, tokens, tokens[:, -1].clone(), cfg, lr=1e-2, weight_decay=1e-2)
        loss, per_sample = tr.step(tokens.to(DEV), torch.zeros(4, device=DEV))
        assert abs(loss.item() - ref_loss) < 2e-5 * max(1, abs(ref_loss)), (step, loss.item(), ref_loss)
        assert abs(per_sample.mean().item() - ref_loss) < 1e-4
    got = m.state_dict()
    for k, v in ref_p.items():
        np.testing.assert_allclose(got[k].cpu().numpy(), v.numpy(), rtol=2e-3, atol=2e-5, err_msg=k)


def test_bf16_training_reduces_loss_with_cuda_graph():
    torch.manual_seed(0)
    m = wm.VqVideoDiffusionModel(data_shape=(4, 8, 8), dim=64, num_classes=32, extents=(1, 2, 2), depth=2, heads=2,
                                 dim_head=32, mlp_dim=64).to(DEV)
    tr = wm.DenoiserTrainer(m, lr=3e-3, use_cuda_graph=True)
    tokens = torch.randint(0, 32, (8, 4, 8, 8), device=DEV)
    r = torch.full((8,), 0.5, device=DEV)
    losses = [tr.step(tokens, r)[0].item() for _ in range(30)]
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.05, losses


C3 = dict(data_shape=(16, 16, 16), dim=256, num_classes=512, extents=(1, 2, 2), depth=4, heads=8, dim_head=32, mlp_dim=256)


def _rel_close(a, b, rtol, atol_frac, what):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    atol = atol_frac * b.abs().max().item() + 1e-7
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert bad.float().mean().item() < 2e-3, (f'{what}: {int(bad.sum())}/{bad.numel()} elements off; max|diff|='
                                               f'{(a - b).abs().max().item():.3e} scale={b.abs().max().item():.3e}')


@pytest.mark.parametrize('prune', [False, True])
def test_bf16_config3_step_matches_oracle(prune):
    """The benchmarked path itself (prune=True: the same through the opt-in receptive-field cone, 5/4/3/2 frames per layer): bf16, BASELINE config-3 shape, fused deferred-bias schedule, CUDA graphs.  One clip,
    r = 0 (no corruption, so that the oracle sees the same inputs): loss and every parameter gradient against the
    fp32 CPU oracle (bf16 bar: rtol 2e-2 plus a scale-relative atol; gradients are sums over 4096 tokens of bf16
    products, so a small fraction of near-zero elements may exceed it)."""
    cfg = O.DenoiserConfig(**C3)
    p = O.init_denoiser_params(cfg, seed=42)
    m = wm.VqVideoDiffusionModel(**C3).to(DEV)
    m.load_state_dict(p)
    m.prune_receptive_field = prune
    tr = wm.DenoiserTrainer(m, lr=1e-4, weight_decay=1e-7, compute_dtype=torch.bfloat16, use_cuda_graph=True)
    tokens = torch.randint(0, 512, (1, 16, 16, 16), generator=torch.Generator().manual_seed(7))
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    ref_loss = O.denoiser_loss(leaves, tokens, tokens[:, -1].clone(), cfg)
    ref_grads = dict(zip(leaves, torch.autograd.grad(ref_loss, list(leaves.values()))))
    loss, per_sample = tr.step(tokens.to(DEV), torch.zeros(1, device=DEV))
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    assert abs(per_sample.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item())
    names = {id(q): n for n, q in m.named_parameters()}
    gsq = 0.0
    for prm, gv in zip(tr._params, tr._grad_views):
        name = names[id(prm)]
        _rel_close(gv, ref_grads[name], 2e-2, 2e-2, name)
        gsq += float(ref_grads[name].double().pow(2).sum())
    # the gradient norm reduced inside the AdamW kernel (main.py:189-193)
    assert abs(tr.grad_norm().item() - gsq ** 0.5) < 3e-2 * gsq ** 0.5


def test_accumulation_steps_equal_one_big_batch_fp32():
    """main.py:205,221-280: two micro-batches of 2 with accumulation_steps=2 == one batch of 4 (exact mode, graphs)."""
    kw = dict(data_shape=(3, 8, 8), dim=32, num_classes=20, extents=(1, 2, 1), depth=2, heads=2, dim_head=16, mlp_dim=40)
    p = O.init_denoiser_params(O.DenoiserConfig(**kw), seed=3)
    tokens = torch.randint(0, 20, (4, 3, 8, 8), generator=torch.Generator().manual_seed(0)).to(DEV)
    r = torch.zeros(4, device=DEV)
    out = []
    for acc in (1, 2):
        m = wm.VqVideoDiffusionModel(**kw).to(DEV)
        m.load_state_dict(p)
        tr = wm.DenoiserTrainer(m, lr=1e-2, weight_decay=1e-2, compute_dtype=torch.float32, use_cuda_graph=True,
                                accumulation_steps=acc, lr_schedule=lambda s: 1e-2 * (s + 1))
        for _ in range(2):
            if acc == 1:
                tr.step(tokens, r)
            else:
                tr.step(tokens[:2], r[:2])
                tr.step(tokens[2:], r[2:])
        out.append({k: v.clone() for k, v in m.state_dict().items()})
        assert abs(tr.dyn[1].item() - 2e-2) < 1e-9            # lr_schedule(step 1) was applied before the second update
    for k in out[0]:
        torch.testing.assert_close(out[0][k], out[1][k], rtol=2e-4, atol=2e-6, msg=k)


def test_trainer_checkpoint_round_trip_and_stock_adamw():
    """DenoiserTrainer.checkpoint() -> a fresh trainer's load_checkpoint() -> identical next step; and the saved
    optimizer_state_dict loads into a stock torch.optim.AdamW over a plain fp32 copy of the model (the reference's
    on-disk format, main.py:297-309).  mlp_dim = 44 and 5-wide heads make parameter sizes that are not multiples of 8
    (flat-buffer alignment)."""
    kw = dict(data_shape=(3, 8, 8), dim=24, num_classes=19, extents=(1, 1, 1), depth=2, heads=3, dim_head=8, mlp_dim=44)
    torch.manual_seed(0)
    m1 = wm.VqVideoDiffusionModel(**kw).to(DEV)
    tr1 = wm.DenoiserTrainer(m1, lr=3e-3, compute_dtype=torch.float32, use_cuda_graph=False)
    g = torch.Generator().manual_seed(1)
    batches = [torch.randint(0, 19, (2, 3, 8, 8), generator=g).to(DEV) for _ in range(4)]
    r = torch.zeros(2, device=DEV)
    for b in batches[:3]:
        tr1.step(b, r)
    ck = tr1.checkpoint(step=3, opt={'note': 'x'}, ema_model_state={'k': torch.ones(1)})
    assert set(ck) == {'step', 'lr', 'model_state_dict', 'ema_model_state_dict', 'optimizer_state_dict', 'opt'}
    assert ck['ema_model_state_dict'] is not None and ck['step'] == 3
    m2 = wm.VqVideoDiffusionModel(**kw).to(DEV)
    tr2 = wm.DenoiserTrainer(m2, lr=1.0, compute_dtype=torch.float32, use_cuda_graph=False)
    assert tr2.load_checkpoint(ck) == 3
    l1, _ = tr1.step(batches[3], r)
    l2, _ = tr2.step(batches[3], r)
    assert abs(l1.item() - l2.item()) < 1e-6
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=k)
    # stock AdamW: continue one step from the checkpoint on the oracle's CPU modules and compare with the trainer
    cfg = O.DenoiserConfig(**kw)
    params = {k: v.clone().requires_grad_(True) for k, v in ck['model_state_dict'].items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1.0, weight_decay=1e-7)
    sd = ck['optimizer_state_dict']
    opt.load_state_dict(sd)
    assert abs(opt.param_groups[0]['lr'] - 3e-3) < 1e-9
    tok = batches[3].cpu()
    O.denoiser_loss(params, tok, tok[:, -1].clone(), cfg).backward()
    opt.step()
    got = m1.state_dict()
    for k, v in params.items():
        torch.testing.assert_close(got[k].cpu(), v.detach(), rtol=2e-4, atol=2e-6, msg=k)


def test_device_loss_aware_sampler_matches_host():
    """LossAwareSamplerEma on the device (wm_loss_hist_update) == the host implementation (importance_sampling.py:35-47)."""
    host = wm.LossAwareSamplerEma(num_histogram_buckets=10, warmup=1, seed=3)
    dev = wm.LossAwareSamplerEma(num_histogram_buckets=10, warmup=1, seed=3, device=DEV)
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        ts, losses = torch.rand(64, generator=g), torch.rand(64, generator=g) * 3
        host.update_with_losses(ts, losses)
        dev.update_with_losses(ts.to(DEV), losses.to(DEV))
    torch.testing.assert_close(dev._weights.cpu(), host._weights, rtol=1e-5, atol=1e-6)
    assert torch.equal(dev._counts.cpu(), host._counts)
    torch.testing.assert_close(dev.weights().cpu(), host.weights(), rtol=1e-5, atol=1e-6)
    r = dev.sample(1000)
    assert r.is_cuda and r.min().item() >= 0 and r.max().item() < 1


def _chi2_same_distribution(a, b, K):
    """Two-sample chi-square statistic / dof between two token samples."""
    ca = torch.bincount(a.reshape(-1).cpu(), minlength=K).double()
    cb = torch.bincount(b.reshape(-1).cpu(), minlength=K).double()
    na, nb = ca.sum(), cb.sum()
    keep = (ca + cb) > 0
    stat = (((ca * (nb / na).sqrt() - cb * (na / nb).sqrt()) ** 2) / (ca + cb))[keep].sum().item()
    return stat / max(1, int(keep.sum()) - 1)


def test_pruned_forward_equals_full_forward_multi_clip():
    """Several clips (the frame slices are strided copies then), short and long cones, fp32 and bf16: logits and
    parameter gradients of the pruned evaluation against the full one on the same device."""
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 3e-2)):
        for S, depth, e_s in ((6, 2, 1), (3, 2, 1), (7, 1, 2), (5, 3, 0)):
            torch.manual_seed(S)
            m = wm.VqVideoDiffusionModel(data_shape=(S, 8, 8), dim=32, num_classes=24, extents=(e_s, 1, 2), depth=depth,
                                         heads=2, dim_head=16, mlp_dim=48).to(DEV).to(dtype)
            tokens = torch.randint(0, 25, (3, S, 8, 8), device=DEV)
            target = torch.randint(0, 24, (3, 8, 8), device=DEV)
            res = []
            for prune in (False, True):
                m.prune_receptive_field = prune
                m.zero_grad(set_to_none=True)
                logits = m(tokens)
                torch.nn.functional.cross_entropy(logits.float().reshape(-1, 24), target.reshape(-1)).backward()
                res.append((logits.detach().float(), {k: q.grad.detach().float() for k, q in m.named_parameters()}))
            scale = res[0][0].abs().max().item()
            assert (res[0][0] - res[1][0]).abs().max().item() <= tol * scale, (dtype, S, depth, e_s)
            for k, g in res[0][1].items():
                gs = g.abs().max().item() + 1e-12
                assert (g - res[1][1][k]).abs().max().item() <= tol * gs, (dtype, S, depth, e_s, k)


@pytest.mark.parametrize('graph,topk', [(False, -1), (True, -1), (True, 3), (False, 3)])
def test_sampler_distribution_matches_oracle(graph, topk):
    """sample_next_frame (eager / captured, with and without top-k) draws from the same distribution as the oracle's
    restatement of main.py:71-111 with identical weights: token histograms of the final frame over many clips
    (chi-square per degree of freedom near 1; a sampler that ignored the logits, the top-k filter or the re-masking
    schedule lands far above the bound), and per-position agreement of the most likely token."""
    kw = dict(data_shape=(3, 4, 4), dim=32, num_classes=12, extents=(1, 1, 1), depth=1, heads=2, dim_head=16, mlp_dim=32)
    cfg = O.DenoiserConfig(**kw)
    p = O.init_denoiser_params(cfg, seed=11)
    p['logit_proj.weight'] = p['logit_proj.weight'] * 6.0            # peaked, context-dependent logits
    m = wm.VqVideoDiffusionModel(**kw).to(DEV).eval()
    m.load_state_dict(p)
    B = 512
    ctx = torch.randint(0, 12, (1, 3, 4, 4), generator=torch.Generator().manual_seed(2)).repeat(B, 1, 1, 1)
    ctx[:, -1] = 12
    ref = O.sample_next_frame(p, ctx.clone(), cfg, iterations=6, gen=torch.Generator().manual_seed(5), sample_topk=topk)
    ours = wm.sample_next_frame(m, ctx.to(DEV), iterations=6, sample_topk=topk, use_cuda_graph=graph, seed=17)
    assert ours.shape == ref.shape and ours.min().item() >= 0 and ours.max().item() < 12
    assert _chi2_same_distribution(ours, ref, 12) < 3.0
    # same context in every clip: per position, the modal token of the two samplers agrees almost everywhere
    mode_o = torch.mode(ours.cpu().reshape(B, -1), dim=0).values
    mode_r = torch.mode(ref.reshape(B, -1), dim=0).values
    assert (mode_o == mode_r).float().mean().item() >= 0.75
    # a uniform sampler is rejected by the same statistic
    uni = torch.randint(0, 12, ref.shape, generator=torch.Generator().manual_seed(1))
    assert _chi2_same_distribution(uni, ref, 12) > 3.0


def test_sample_frames_shifts_context_and_decodes():
    """evaluate_model's outer loop (main.py:62-117): num_steps frames, context shifted by one after each, decoded by the
    VQ auto-encoder; the caller's tokens are untouched."""
    torch.manual_seed(4)
    kw = dict(data_shape=(3, 8, 8), dim=32, num_classes=16, extents=(1, 1, 1), depth=1, heads=2, dim_head=16, mlp_dim=32)
    m = wm.VqVideoDiffusionModel(**kw).to(DEV).eval()
    ae = wm.VqAutoEncoder(8, 16, downscale_steps=2, hidden_planes=16, in_channels=1).to(DEV).eval()
    tokens = torch.randint(0, 16, (4, 3, 8, 8), device=DEV)
    keep = tokens.clone()
    frames, decoded = wm.sample_frames(m, tokens, num_steps=3, iterations=4, decoder=ae, use_cuda_graph=True)
    assert frames.shape == (3, 4, 8, 8) and frames.min().item() >= 0 and frames.max().item() < 16
    assert len(decoded) == 3 and decoded[0].shape == (4, 1, 32, 32)
    assert torch.equal(tokens, keep)
    st = next(iter(m.__dict__['_wm_sample_graphs'].values()))
    # after the last shift the context holds: original frame 2.., then the generated frames (main.py:115)
    assert torch.equal(st.work[:, 0], frames[1]) and torch.equal(st.work[:, 1], frames[2])
    torch.testing.assert_close(decoded[1], ae.decode(frames[1]))


def test_sampler_produces_valid_tokens_and_is_batch_independent():
    torch.manual_seed(1)
    m = wm.VqVideoDiffusionModel(data_shape=(3, 8, 8), dim=32, num_classes=16, extents=(1, 1, 1), depth=1, heads=2,
                                 dim_head=16, mlp_dim=32).to(DEV).eval()
    tokens = torch.randint(0, 16, (4, 3, 8, 8), device=DEV)
    tokens[:, -1] = 16
    out = wm.sample_next_frame(m, tokens, iterations=5)
    assert out.shape == (4, 8, 8) and out.min().item() >= 0 and out.max().item() < 16
    # the denoiser forward is per-clip: a clip's logits do not depend on its batch neighbours,
    # which is what makes batch-sharded multi-GPU sampling exact
    with torch.no_grad():
        a = m(tokens)
        b = torch.cat([m(tokens[:2]), m(tokens[2:])])
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_graphed_sampler_reuses_one_capture():
    """The captured sampling iteration is cached on the model and replayed for later frames (fresh inputs are copied
    into the captured buffers); draws are valid tokens.  With iterations=1 there is no re-masking and the draw comes
    from the zero-initialised logits, i.e. uniform: graph and eager sampling then differ only by the RNG stream."""
    torch.manual_seed(2)
    m = wm.VqVideoDiffusionModel(data_shape=(3, 8, 8), dim=32, num_classes=16, extents=(1, 1, 1), depth=1, heads=2,
                                 dim_head=16, mlp_dim=32).to(DEV).eval()
    tokens = torch.randint(0, 16, (4, 3, 8, 8), device=DEV)
    tokens[:, -1] = 16
    for _ in range(2):
        out = wm.sample_next_frame(m, tokens, iterations=4, use_cuda_graph=True)
        assert out.shape == (4, 8, 8) and out.min().item() >= 0 and out.max().item() < 16
    out2 = wm.sample_next_frame(m, tokens.flip(0), iterations=4, use_cuda_graph=True)
    assert out2.shape == (4, 8, 8) and out2.min().item() >= 0 and out2.max().item() < 16
    assert len(m.__dict__['_wm_sample_graphs']) == 1
    # the frames before the last one are inputs only: the sampler must not have modified the caller's tensor
    assert (tokens[:, -1] == 16).all()


def test_sparse_context_model_against_reference_fixture():
    """minecraft/sparse_diffusion.py:75-111 on the GPU: logits, loss and all parameter gradients of the drop-in
    VqSparseDiffusionModel against the fixture written by the reference's dense Transformer (fp32)."""
    from world_modelz_b200.sparse_diffusion import VqSparseDiffusionModel
    f = load('sparse_small.npz')
    c = [int(v) for v in f['cfg']]
    m = VqSparseDiffusionModel(shape=tuple(c[0:3]), dim=c[3], num_classes=c[4], depth=c[5], dim_head=c[6], mlp_dim=c[7],
                               heads=c[8]).to(DEV)
    m.load_state_dict(state_dict_of(f))
    logits = m(torch.from_numpy(f['tokens']).to(DEV), torch.from_numpy(f['indices']).to(DEV))
    np.testing.assert_allclose(logits.detach().cpu().numpy(), f['logits'], rtol=1e-4, atol=2e-5)
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, c[4]), torch.from_numpy(f['target']).to(DEV).reshape(-1))
    assert abs(loss.item() - float(f['loss'])) < 1e-5
    loss.backward()
    for k, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), f['grad/' + k], rtol=2e-4, atol=5e-6, err_msg=k)
