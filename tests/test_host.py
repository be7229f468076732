"""Host-side logic on CPU: drop-in module surface, corruption statistics, sampler,
batch sharding and the world_size-2 gradient exchange over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import world_modelz_b200 as wm
from world_modelz_b200 import parallel
from world_modelz_b200.denoiser import LossAwareSamplerEma
from oracle import local3d as O
from tests._golden import load, state_dict_of


def test_state_dict_keys_match_reference():
    f = load('denoiser_small.npz')
    ref = state_dict_of(f)
    m = wm.VqVideoDiffusionModel(data_shape=(4, 6, 6), dim=32, num_classes=17, extents=(1, 1, 2), depth=2, heads=2,
                                 dim_head=16, mlp_dim=48)
    sd = m.state_dict()
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref)            # strict load of a reference checkpoint


def test_attention_module_surface():
    f = load('attn_small.npz')
    m = wm.Local3dAttention((1, 2, 2), 32, heads=2, dim_head=16)
    m.load_state_dict(state_dict_of(f))
    assert m.heads == 2 and m.scale == 16 ** -0.5 and tuple(m.extents) == (1, 2, 2) and m.use_checkpointing
    f2 = load('attn_noproj.npz')
    m2 = wm.Local3dAttention((2, 1, 1), 16, heads=1, dim_head=16)
    assert isinstance(m2.to_out, torch.nn.Identity)
    m2.load_state_dict(state_dict_of(f2))


def test_vq_module_surface():
    vq = wm.VectorQuantizerEMA(64, 512)
    assert set(vq.state_dict()) == {'embedding', 'cluster_size'}          # vq.py:16-20 persistence flags
    assert vq.embedding.shape == (1, 512, 64) and vq.latent_offsets.shape == (1, 1)
    assert vq.simple_update is False and vq.laplace_smoothing is True
    idx = torch.tensor([[3], [500]])
    assert torch.equal(vq.decode(idx)[:, 0], vq.embedding[0, [3, 500]])
    vq2 = wm.VectorQuantizerEMA(8, 12, num_latents=3)
    f = load('vq_multilatent.npz')
    vq2.embedding.copy_(torch.from_numpy(f['embedding1']))
    got = vq2.decode(torch.from_numpy(f['idx']).view(5, 7, 3))
    assert np.array_equal(got.numpy(), f['decode'])


def test_corruption_matches_reference_distribution():
    K, B = 16, 4000
    tokens = torch.randint(0, K, (B, 2, 2, 2))
    r = torch.full((B,), 0.6)
    torch.manual_seed(0)
    ours, tgt = wm.corrupt_last_frame(tokens, r, K)
    theirs, tgt2 = O.corrupt_last_frame(tokens, r, K, gen=torch.Generator().manual_seed(1))
    assert torch.equal(tgt, tokens[:, -1]) and torch.equal(tgt2, tgt)
    assert torch.equal(ours[:, :-1], tokens[:, :-1])
    for out in (ours, theirs):
        masked = (out[:, -1] == K).float().mean().item()
        assert abs(masked - 0.6) < 0.02
        keep = out[:, -1] != K
        changed = (out[:, -1][keep] != tokens[:, -1][keep]).float().mean().item()
        assert abs(changed - 0.06 * (K - 1) / K) < 0.01          # r * 0.1 * (1 - 1/K)


def test_loss_aware_sampler_semantics():
    s = LossAwareSamplerEma(num_histogram_buckets=10, warmup=1, seed=3)
    assert not s.warmed_up() and torch.equal(s.weights(), torch.ones(10))
    ts = torch.tensor([0.05, 0.05, 0.95])
    s.update_with_losses(ts, torch.tensor([2.0, 4.0, 1.0]))
    # two sequential EMA updates on bucket 0 (importance_sampling.py:40-41)
    assert abs(s._weights[0].item() - ((1 * 0.9 + 2 * 0.1) * 0.9 + 4 * 0.1)) < 1e-6
    assert abs(s._weights[9].item() - (0.9 + 0.1)) < 1e-6
    for _ in range(3):
        s.update_with_losses(torch.arange(10) / 10 + 0.01, torch.ones(10))
    assert s.warmed_up()
    w = s.weights()
    assert abs(w.sum().item() - 1) < 1e-5
    r = s.sample(1000)
    assert r.min() >= 0 and r.max() < 1


def test_shard_range_covers_everything():
    for total in (64, 7, 1, 0):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, _, w = parallel.init_from_env('gloo')
    cfg = O.DenoiserConfig(data_shape=(3, 4, 4), dim=16, num_classes=9, extents=(1, 1, 1), depth=1, heads=2,
                           dim_head=8, mlp_dim=24)
    p = O.init_denoiser_params(cfg, seed=5)
    g = torch.Generator().manual_seed(11)
    tokens = torch.randint(0, 10, (4, 3, 4, 4), generator=g)
    target = torch.randint(0, 9, (4, 4, 4), generator=g)
    b, e = parallel.shard_range(4, r, w)
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    O.denoiser_loss(leaves, tokens[b:e], target[b:e], cfg).backward()
    flat = torch.cat([v.grad.reshape(-1) for v in leaves.values()])
    parallel.allreduce_mean_(flat)
    if r == 0:
        full = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        O.denoiser_loss(full, tokens, target, cfg).backward()
        ref = torch.cat([v.grad.reshape(-1) for v in full.values()])
        out.put(float((flat - ref).abs().max() / ref.abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_gloo():
    """world_size=2: averaging per-shard gradients reproduces the full-batch gradient."""
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() < 1e-5


def test_deferred_bias_host_paths_on_cpu():
    """Host logic of the deferred-bias schedule that needs no GPU: widths the fused kernel does not tile go through
    stock ops with the same math, and modules hand back ``bias=None`` when they cannot defer (CPU tensors)."""
    import torch
    import torch.nn.functional as F
    from world_modelz_b200 import ops
    from world_modelz_b200.local_3d_attention import FeedForward
    torch.manual_seed(0)
    res, delta = torch.randn(7, 12), torch.randn(7, 12)            # dim 12: not a multiple of 8 -> stock-op path
    bias, gamma, beta = torch.randn(12), torch.rand(12) + 0.5, torch.randn(12)
    total, y = ops.add_layernorm(res, delta, gamma, beta, 1e-5, bias)
    torch.testing.assert_close(total, res + delta + bias)
    torch.testing.assert_close(y, F.layer_norm(res + delta + bias, (12,), gamma, beta, 1e-5))
    total2, _ = ops.add_layernorm(res, None, gamma, beta, 1e-5)
    assert total2 is res
    ff = FeedForward(12, 24)
    x = torch.randn(3, 12)
    out, b = ff.forward_deferred_bias(x)
    assert b is None
    torch.testing.assert_close(out, ff(x))
    torch.testing.assert_close(ops.bias_gelu(x, torch.zeros(12)), F.gelu(x))     # CPU tensors: stock ops


def test_reference_checkpoint_format_round_trip():
    """``checkpoint.py``: flat AdamW moments <-> ``torch.optim.AdamW.state_dict()`` (the reference's
    ``optimizer_state_dict``, main.py:300-307), checked against a real AdamW on the CPU."""
    import torch
    from world_modelz_b200 import checkpoint as ck
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4, weight_decay=1e-7)
    for _ in range(3):
        opt.zero_grad()
        net(torch.randn(4, 6)).square().mean().backward()
        opt.step()
    ref_state = opt.state_dict()
    params = list(net.parameters())
    shapes = ck.param_shapes(params)
    n = sum(p.numel() for p in params)
    m, v = torch.full((n,), 7.0), torch.full((n,), 7.0)
    steps = ck.adamw_state_to_flat(ref_state, shapes, m, v)
    assert steps == 3
    off = 0
    for i, p in enumerate(params):
        torch.testing.assert_close(m[off:off + p.numel()].view_as(p), ref_state['state'][i]['exp_avg'])
        torch.testing.assert_close(v[off:off + p.numel()].view_as(p), ref_state['state'][i]['exp_avg_sq'])
        off += p.numel()
    back = ck.flat_to_adamw_state(shapes, m, v, steps, lr=3e-4, weight_decay=1e-7)
    opt2 = torch.optim.AdamW(net.parameters(), lr=1.0)
    opt2.load_state_dict(back)                                   # a stock AdamW accepts what we write
    for i in range(len(params)):
        torch.testing.assert_close(opt2.state_dict()['state'][i]['exp_avg'], ref_state['state'][i]['exp_avg'])
        assert float(opt2.state_dict()['state'][i]['step']) == 3.0
    assert opt2.state_dict()['param_groups'][0]['lr'] == 3e-4
    data = ck.make_checkpoint(net.state_dict(), back, step=3, lr=3e-4, opt={'dim': 6})
    assert set(data) == {'step', 'lr', 'model_state_dict', 'ema_model_state_dict', 'optimizer_state_dict', 'opt'}
    # an optimizer that has not stepped yet: empty state, zeroed moments
    fresh = torch.optim.AdamW(net.parameters()).state_dict()
    assert ck.adamw_state_to_flat(fresh, shapes, m, v) == 0 and float(m.abs().sum()) == 0.0
    import pytest
    with pytest.raises(ValueError):
        ck.adamw_state_to_flat(ref_state, shapes[:-1], m, v)


def test_bench_stdout_guard_keeps_native_banners_off_stdout(tmp_path):
    """bench.py prints ONE JSON line on stdout; anything a native library writes to fd 1 before that goes to stderr."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'guard.py'
    script.write_text(
        'import os, sys, json, importlib.util\n'
        f'spec = importlib.util.spec_from_file_location("bench", r"{root}/bench.py")\n'
        'b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n'
        'g = b.StdoutGuard()\n'
        'os.write(1, b"NCCL version banner\\n")\n'
        'print("guarded python print")\n'
        'g.release()\n'
        'print(json.dumps({"ok": 1}), flush=True)\n')
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip() == '{"ok": 1}'
    assert 'NCCL version banner' in res.stderr and 'guarded python print' in res.stderr
