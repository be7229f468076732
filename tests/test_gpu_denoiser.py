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


def test_denoiser_against_reference_fixture_fp32():
    f = load('denoiser_small.npz')
    kw, _ = _small(f)
    m = wm.VqVideoDiffusionModel(**kw).to(DEV)
    m.load_state_dict(state_dict_of(f))
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
    for it in (4, 30):
        out = wm.sample_next_frame(m, tokens, iterations=it, use_cuda_graph=True)
        assert out.shape == (4, 8, 8) and out.min().item() >= 0 and out.max().item() < 16
    out2 = wm.sample_next_frame(m, tokens.flip(0), iterations=4, use_cuda_graph=True)
    assert out2.shape == (4, 8, 8) and out2.min().item() >= 0 and out2.max().item() < 16
    assert len(m.__dict__['_wm_sample_graphs']) == 1
    # the frames before the last one are inputs only: the sampler must not have modified the caller's tensor
    assert (tokens[:, -1] == 16).all()
