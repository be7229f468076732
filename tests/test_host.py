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


def test_modules_refuse_cpu_tensors():
    """north_star: no CPU fallback.  The drop-in modules raise on CPU tensors instead of quietly running stock ops."""
    from world_modelz_b200.local_3d_attention import FeedForward
    ff = FeedForward(16, 24)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ff(torch.randn(3, 16))
    m = wm.VqVideoDiffusionModel(data_shape=(2, 4, 4), dim=16, num_classes=8, extents=(1, 1, 1), depth=1, heads=2,
                                 dim_head=8, mlp_dim=16)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.zeros(1, 2, 4, 4, dtype=torch.long))
    vq = wm.VectorQuantizerEMA(8, 16)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        vq(torch.randn(2, 4, 8))


def test_vqautoencoder_surface_and_conv_stacks_match_reference_fixture():
    """``VqAutoEncoder`` (train_vqae.py:22-55): state_dict keys / shapes equal the reference's, a strict load of its
    checkpoint works, and the (stock-op) encoder / decoder stacks reproduce the reference's latents and decoded
    frames on the CPU; the quantizer step in between needs the GPU (tests/test_gpu_vq.py)."""
    f = load('vqae_small.npz')
    emb, K, steps, hidden, cin = (int(v) for v in f['cfg'])
    ae = wm.VqAutoEncoder(emb, K, downscale_steps=steps, hidden_planes=hidden, in_channels=cin)
    ref = state_dict_of(f)
    sd = ae.state_dict()
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    ae.load_state_dict(ref)
    ae.eval()
    with torch.no_grad():
        h = ae.encoder(torch.from_numpy(f['frames'])).permute(0, 2, 3, 1)
        np.testing.assert_allclose(h.numpy(), f['latents'], rtol=1e-5, atol=1e-6)
        dec = ae.decoder(ae.vq.decode(torch.from_numpy(f['z'])).permute(0, 3, 1, 2))
        np.testing.assert_allclose(dec.numpy(), f['decoded'], rtol=1e-5, atol=1e-6)


def test_oracle_sampler_topk_and_frame_shift():
    """Oracle side of the multi-frame sampler (main.py:39-43, 62-117): top-k keeps exactly the k largest logits
    (ties at the k-th value stay), frames shift by one after each generated frame."""
    lg = torch.tensor([[0.1, 3.0, 2.0, 2.0, -1.0]])
    out = O.top_k_logits(lg, 2)
    assert torch.isinf(out[0, [0, 4]]).all() and torch.equal(out[0, 1:4], lg[0, 1:4])
    cfg = O.DenoiserConfig(data_shape=(3, 4, 4), dim=16, num_classes=8, extents=(1, 1, 1), depth=1, heads=2,
                           dim_head=8, mlp_dim=16)
    p = O.init_denoiser_params(cfg, seed=1)
    tok = torch.randint(0, 8, (2, 3, 4, 4), generator=torch.Generator().manual_seed(0))
    frames = O.sample_frames(p, tok, cfg, 2, iterations=3, gen=torch.Generator().manual_seed(0), sample_topk=3)
    assert frames.shape == (2, 2, 4, 4) and frames.min() >= 0 and frames.max() < 8
    assert torch.equal(tok, tok.clone())          # the caller's tensor is not touched


def test_flat_buffer_offsets_are_aligned_in_checkpoint_mapping():
    """Parameters whose sizes are not multiples of 8 must not misalign their successors in the trainer's flat buffers:
    the checkpoint mapping takes explicit (aligned) offsets."""
    from world_modelz_b200 import checkpoint as ck
    from world_modelz_b200.denoiser import _align
    shapes = [torch.Size([3, 5]), torch.Size([7]), torch.Size([4, 4])]
    offs, off = [], 0
    for shp in shapes:
        offs.append(off)
        off = _align(off + shp.numel())
    assert offs == [0, 16, 24] and off == 40
    m, v = torch.arange(40.), torch.arange(40.) * 2
    st = ck.flat_to_adamw_state(shapes, m, v, 3, lr=1e-3, offsets=offs)
    assert torch.equal(st['state'][1]['exp_avg'], torch.arange(16., 23.))
    m2, v2 = torch.zeros(40), torch.zeros(40)
    assert ck.adamw_state_to_flat(st, shapes, m2, v2, offsets=offs) == 3
    assert torch.equal(m2[24:40], m[24:40]) and torch.equal(v2[16:23], v[16:23]) and m2[15] == 0


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


def test_sparse_model_surface_and_position_samplers():
    """VqSparseDiffusionModel keeps the reference's state_dict keys (strict load of its checkpoint); the position
    samplers (sparse_diffusion.py:31-72) return unique in-window positions without per-sample loops."""
    from world_modelz_b200.sparse_diffusion import VqSparseDiffusionModel, sample_flat_positions, sample_time_dependent
    f = load('sparse_small.npz')
    c = [int(v) for v in f['cfg']]
    m = VqSparseDiffusionModel(shape=tuple(c[0:3]), dim=c[3], num_classes=c[4], depth=c[5], dim_head=c[6], mlp_dim=c[7],
                               heads=c[8])
    ref = state_dict_of(f)
    assert set(m.state_dict()) == set(ref)
    m.load_state_dict(ref)
    torch.manual_seed(0)
    p = sample_flat_positions(5, 100, 8, 6, 6, 'cpu')
    assert p.shape == (5, 100) and p.min() >= 0 and p.max() < 288
    flat = p.reshape(-1)
    assert flat[:288].unique().numel() == 288                      # the first permutation-sized run has no repeats
    t = torch.tensor([0.0, 0.3, 1.0, 0.6])
    q = sample_time_dependent(4, 50, 8, 6, 6, t, 'cpu', o=torch.tensor([0.0, 0.99, 0.5, 0.2]))
    assert q.shape == (4, 50)
    for i in range(4):
        assert q[i].unique().numel() == 50                         # without replacement
    # t = 0: the window is min_sample_window = ceil(50 / 36) = 2 frames, starting at frame 0 (o = 0)
    assert q[0].max() < 2 * 36
    # window length grows with t (sparse_diffusion.py:58-59): floor(2 + t * 7), clamped to 6 frames
    frames = [(q[i].div(36, rounding_mode='trunc')).unique().numel() for i in range(4)]
    assert frames[0] <= 2 and frames[1] <= 4 and frames[2] <= 6


def test_last_frame_cone_is_the_reference_models_dependency_cone():
    """Local3dAttentionTransformer.last_frame_cone against the reference's semantics (the oracle's CPU denoiser): tokens
    in frames below cone[0] cannot change the last frame's logits, a token in frame cone[0] can."""
    import world_modelz_b200 as wm
    from oracle import local3d as O
    for S, depth, e_s in ((7, 2, 1), (6, 1, 2), (4, 3, 1)):
        kw = dict(data_shape=(S, 4, 4), dim=16, num_classes=10, extents=(e_s, 1, 1), depth=depth, heads=2, dim_head=8, mlp_dim=16)
        cfg = O.DenoiserConfig(**kw)
        p = O.init_denoiser_params(cfg, seed=S)
        cone = wm.VqVideoDiffusionModel(**kw).transformer.last_frame_cone(S)
        assert cone[-1] == S - 1 and cone[0] == max(0, S - 1 - depth * e_s) and all(b - a in (0, e_s) or a == 0 for a, b in zip(cone, cone[1:]))
        tokens = torch.randint(0, 10, (1, S, 4, 4), generator=torch.Generator().manual_seed(0))
        base = O.denoiser_forward(p, tokens, cfg)
        if cone[0] > 0:
            dead = tokens.clone()
            dead[:, :cone[0]] = (dead[:, :cone[0]] + 3) % 10
            assert torch.equal(O.denoiser_forward(p, dead, cfg), base)
        live = tokens.clone()
        live[:, cone[0]] = (live[:, cone[0]] + 3) % 10
        assert not torch.equal(O.denoiser_forward(p, live, cfg), base)


def test_runner_up_from_two_partitions_is_exact():
    """The argument behind the VQ scanners (csrc/vq_tc.cu): keys that are unique within a 16-key chunk (the position sits
    in their low 4 bits) are covered by two partitions, the chunks and the 16 position classes; then
    min(second smallest chunk minimum, second smallest class minimum) is the second smallest key.  Checked in numpy on
    random key sets, with planted near-ties in the same chunk, at the same position of another chunk, and exact duplicates
    of the (score, position) pair in another chunk."""
    rng = np.random.default_rng(0)
    for trial in range(400):
        nchunk = int(rng.integers(1, 17))
        score = rng.integers(0, 1 << 20, size=(nchunk, 16), dtype=np.int64)
        if trial % 4 == 1 and nchunk > 1:                        # runner-up at the winner's position in another chunk
            c, p = np.unravel_index(np.argmin(score), score.shape)
            score[(c + 1) % nchunk, p] = score[c, p] + (trial % 3)
        if trial % 4 == 2:                                       # runner-up in the winner's own chunk
            c, p = np.unravel_index(np.argmin(score), score.shape)
            score[c, (p + 5) % 16] = score[c, p] + (trial % 2)
        keys = score * 16 + np.arange(16)[None, :]               # (bits - bits(1.0)) * 16 + position
        flat = np.sort(keys.reshape(-1))
        truth = flat[1] if flat.size > 1 else None
        chunk_min = np.sort(keys.min(axis=1))
        class_min = keys.min(axis=0)
        m1 = chunk_min[0]
        second_chunk = chunk_min[1] if nchunk > 1 else np.iinfo(np.int64).max
        wrapped = (class_min - (m1 + 1)) % (1 << 40)             # the winner's class wraps to the top, as in the kernel
        second_class = wrapped.min() + m1 + 1
        assert min(second_chunk, second_class) == truth, (trial, nchunk)
