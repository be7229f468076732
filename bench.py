#!/usr/bin/env python
"""Benchmark of the local-3D-attention / VQ denoiser hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path: the unmodified modules from
                                                             # /root/reference when that tree exists, else the oracle port

A *step* is one denoiser training step (corruption, forward, CE, backward, gradient
all-reduce, AdamW) over one batch of synthetic clips of BASELINE config 3: 16 frames x
16x16 VQ tokens, dim 256, 8 heads x 32, window 3x5x5, depth 4, mlp 256, K=512, bf16.
Weak scaling: every GPU gets `--clips-per-gpu` clips; `value` is whole-job clips/s.
The same JSON line carries the attention-core fwd+bwd tokens/s (the other half of the
metric), its roofline, the VQ kernel's latents/s and a CPU baseline from the oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

C3 = dict(data_shape=(16, 16, 16), dim=256, num_classes=512, extents=(1, 2, 2), depth=4, heads=8, dim_head=32,
          mlp_dim=256)
METRIC = 'denoiser_train_clips_per_s'
UNIT = 'clips/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--clips-per-gpu', type=int, default=32)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-micro', action='store_true', help='skip the attention / VQ kernel timings')
    return ap.parse_args()


def config_dict(n_gpus, clips_per_gpu):
    return {'workload': 'config3: denoiser train step, 16x16x16-token clips, dim 256, 8 heads x 32, window 3x5x5, '
                        'depth 4, mlp 256, K=512',
            'global_batch': n_gpus * clips_per_gpu, 'clips_per_gpu': clips_per_gpu, 'tokens_per_clip': 4096,
            'parallelism': f'dp{n_gpus}', 'l2': 'working set per step (activations >1 GB) exceeds the 126 MB L2'}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get('hbm_gbs', 6650.0), p.get('bf16_tflops_sustained', 1400.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1400.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------ CPU reference arm
REF_DIR = '/root/reference/vq-video-diffusion'
# measured in the build container (8-core Xeon, this commit): seconds per 1-clip config-3 train step
PORT_VS_REFERENCE = 'oracle port 2.2 s/step vs unmodified reference modules 5.7 s/step on the 8-core build box'


def _reference_step_fn(clips):
    """One config-3 training step on the UNMODIFIED reference modules (local_3d_attention.py + a Linear head + AdamW, which
    is all main.py:25-36,266-283,433 does), imported read-only from /root/reference.  None when that tree is absent
    (it does not exist on the GPU box)."""
    if not os.path.isdir(REF_DIR):
        return None
    try:
        sys.path.insert(0, REF_DIR)
        from local_3d_attention import Local3dAttentionTransformer   # noqa
    except Exception:
        return None
    finally:
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)
    torch.manual_seed(42)
    kw = dict(C3)
    kw['num_classes'] = C3['num_classes'] + 1
    tr = Local3dAttentionTransformer(**kw)
    head = torch.nn.Linear(C3['dim'], C3['num_classes'])
    opt = torch.optim.AdamW(list(tr.parameters()) + list(head.parameters()), lr=1e-4, weight_decay=1e-7)
    from oracle import local3d as O

    def step(tokens, r, gen):
        corrupted, target = O.corrupt_last_frame(tokens, r, C3['num_classes'], gen=gen)
        opt.zero_grad()
        logits = head(tr(corrupted)[:, -1])
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, C3['num_classes']), target.reshape(-1))
        loss.backward()
        opt.step()
    return step


def cpu_train_sample(steps, warmup, clips, allow_reference=True):
    """Reference CPU path on `clips` clips per step, fp32, all host threads: the reference's own modules when
    /root/reference is importable, else the oracle's train_step.  Returns (seconds, steps, kind)."""
    from oracle import local3d as O
    torch.set_num_threads(os.cpu_count() or 1)
    ref_step = _reference_step_fn(clips) if allow_reference else None
    cfg = O.DenoiserConfig(**C3)
    p = O.init_denoiser_params(cfg, seed=42)
    state = {}
    g = torch.Generator().manual_seed(42)
    times = []
    for it in range(warmup + steps):
        tokens = torch.randint(0, cfg.num_classes, (clips, *cfg.data_shape), generator=g)
        r = torch.rand(clips, generator=g)
        t0 = time.perf_counter()
        if ref_step is not None:
            ref_step(tokens, r, g)
        else:
            corrupted, target = O.corrupt_last_frame(tokens, r, cfg.num_classes, gen=g)
            O.train_step(p, state, it + 1, corrupted, target, cfg)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times), len(times), ('reference' if ref_step is not None else 'port')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    clips = 1
    steps = min(args.steps, 40)                  # bounded sample: the whole run must end within a few minutes
    total, n, kind = cpu_train_sample(steps, min(args.warmup, 1), clips)
    value = clips * n / total
    cores = os.cpu_count() or 1
    what = 'the unmodified reference modules (/root/reference)' if kind == 'reference' else \
        f'the oracle port of the reference modules ({PORT_VS_REFERENCE})'
    sample = f'{n} steps x {clips} clip(s) of config 3, fp32, {what} on {cores} host threads'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': n,
            'warmup': min(args.warmup, 1), 'ms_per_step': 1e3 * total / n, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic', 'config': config_dict(args.gpus, clips),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------- kernel timings
def time_kernel(fn, iters, stream=None):
    """Average device time of fn() in ms over `iters` calls (CUDA events on the current stream)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def micro_benchmarks(dev, clips, hbm_gbs, tf_peak):
    """Attention core fwd / bwd on the config-3 shape and the VQ search on config-2 widths."""
    from world_modelz_b200 import ops
    out = {}
    S, H, W = C3['data_shape']
    heads, d, ext = C3['heads'], C3['dim_head'], C3['extents']
    inner = heads * d
    g = torch.Generator(device=dev).manual_seed(1)
    q, k, v, do = (torch.randn(clips, S, H, W, inner, device=dev, generator=g).bfloat16() for _ in range(4))
    scale = d ** -0.5
    o, lse = ops.attn_forward(q, k, v, heads, ext, scale)
    t_f = time_kernel(lambda: ops.attn_forward(q, k, v, heads, ext, scale), 20)
    t_b = time_kernel(lambda: ops.attn_backward(q, k, v, o, lse, do, heads, ext, scale), 10)
    tokens = clips * S * H * W
    wn = (2 * ext[0] + 1) * (2 * ext[1] + 1) * (2 * ext[2] + 1)
    bytes_f = tokens * (4 * inner * 2 + heads * 4)
    bytes_fb = tokens * (12 * inner * 2 + 3 * heads * 4)
    flops_f = 4.0 * wn * inner * tokens
    out['attn'] = {'shape': f'{clips}x{S}x{H}x{W}x({heads}x{d}) window {wn}', 'fwd_ms': t_f, 'bwd_ms': t_b,
                   'fwd_bwd_tokens_per_s': tokens / ((t_f + t_b) * 1e-3), 'fwd_tokens_per_s': tokens / (t_f * 1e-3),
                   'fwd_algorithmic_gbs': bytes_f / (t_f * 1e-3) / 1e9,
                   'fwd_bwd_algorithmic_gbs': bytes_fb / ((t_f + t_b) * 1e-3) / 1e9,
                   'fwd_algorithmic_tflops': flops_f / (t_f * 1e-3) / 1e12,
                   'fwd_bwd_algorithmic_tflops': 3.5 * flops_f / ((t_f + t_b) * 1e-3) / 1e12,
                   'tensor_cores': ops.uses_tensor_cores(S, H, W, heads, d, ext)}
    # roofline of the dominant kernel of the path: attention forward+backward, HBM-bound algorithmically
    ach = bytes_fb / ((t_f + t_b) * 1e-3) / 1e9
    out['roofline'] = {'kernel': 'local-3D attention core fwd+bwd (l3d_fwd_tc + fix-up scan, delta, dQ, dK/dV)', 'bound': 'hbm', 'achieved': ach,
                       'peak': hbm_gbs, 'unit': 'GB/s', 'frac': ach / hbm_gbs, 'traffic': None,
                       'algorithmic_bytes_per_launch_set': bytes_fb,
                       'note': 'algorithmic bytes = 12*inner*2 B + 12*heads B per token (q,k,v,o,dO read; o,dq,dk,dv written)'}
    # config 4 widths: 32x32x32 tokens, 4 heads x 128, window 5x7x7 (2 clips)
    S4, heads4, d4, ext4 = 32, 4, 128, (2, 3, 3)
    q4, k4, v4, do4 = (torch.randn(2, S4, S4, S4, heads4 * d4, device=dev, generator=g).bfloat16() for _ in range(4))
    o4, lse4 = ops.attn_forward(q4, k4, v4, heads4, ext4, d4 ** -0.5)
    t4f = time_kernel(lambda: ops.attn_forward(q4, k4, v4, heads4, ext4, d4 ** -0.5), 10)
    t4b = time_kernel(lambda: ops.attn_backward(q4, k4, v4, o4, lse4, do4, heads4, ext4, d4 ** -0.5), 5)
    tok4 = 2 * S4 ** 3
    out['attn_config4'] = {'shape': '2x32x32x32x(4x128) window 245', 'fwd_ms': t4f, 'bwd_ms': t4b,
                           'fwd_bwd_tokens_per_s': tok4 / ((t4f + t4b) * 1e-3),
                           'fwd_bwd_algorithmic_tflops': 3.5 * 4.0 * 245 * 512 * tok4 / ((t4f + t4b) * 1e-3) / 1e12,
                           'fwd_bwd_algorithmic_gbs': tok4 * (12 * 512 * 2 + 12 * 4) / ((t4f + t4b) * 1e-3) / 1e9}
    del q4, k4, v4, do4, o4, lse4
    # VQ nearest: 4096 frames of 16x16 latents (D=64) against 512 codes
    n = 1 << 20
    x = torch.randn(n, 1, 64, device=dev, generator=g)
    cb = torch.randn(1, 512, 64, device=dev, generator=g)
    t_v = time_kernel(lambda: ops.vq_nearest(x, cb), 5)
    out['vq'] = {'shape': f'{n} latents x 64 vs 512 codes (fp32, bit-exact indices)', 'ms': t_v,
                 'latents_per_s': n / (t_v * 1e-3), 'algorithmic_gbs': n * 524 / (t_v * 1e-3) / 1e9,
                 'kernel': 'split-bf16 tcgen05 filter (3 products + norm step) + vq_settle_kernel (exact settlement of undecided rows)'}
    ach_v = n * 524 / (t_v * 1e-3) / 1e9
    out['roofline_vq'] = {'kernel': 'vq_nearest_tc_kernel + vq_settle_kernel', 'bound': 'hbm', 'achieved': ach_v, 'peak': hbm_gbs, 'unit': 'GB/s',
                          'frac': ach_v / hbm_gbs, 'traffic': None,
                          'note': 'algorithmic bytes = 4D read + 4D straight-through value + 8 index + 4 error = 524 B per latent (D = 64)'}
    return out


def config4_transformer_benchmark(dev, tf_peak, clips=1, iters=3):
    """BASELINE config 4 (X1): Local3dAttentionTransformer at the sparse_diffusion sizing -- 32x32x32 tokens, dim 512,
    4 heads x 128, depth 8, mlp 1024, window 5x7x7 (minecraft/sparse_diffusion.py:233,250-253,362 on
    minecraft/main2.py:20,31) -- forward + backward of a mean loss, bf16, `clips` clip(s)."""
    import world_modelz_b200 as wm
    torch.manual_seed(4)
    m = wm.Local3dAttentionTransformer(data_shape=(32, 32, 32), dim=512, num_classes=1024, extents=(2, 3, 3), depth=8,
                                       heads=4, dim_head=128, mlp_dim=1024).to(dev).bfloat16()
    tokens = torch.randint(0, 1024, (clips, 32, 32, 32), device=dev)

    def step():
        for p in m.parameters():
            p.grad = None
        m(tokens).float().mean().backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # per token-layer: 4 projections + 2 MLP GEMMs (2 flop/MAC) + attention core 4*Wn*inner; x3 for fwd+bwd
    flop_clip = 3.0 * 32768 * 8 * (2 * (4 * 512 * 512 + 2 * 512 * 1024) + 4 * 245 * 512)
    del m
    return {'shape': f'{clips} x 32x32x32 tokens, dim 512, 4 heads x 128, depth 8, mlp 1024, window 245', 'fwd_bwd_ms': ms,
            'clips_per_s': clips / (ms * 1e-3), 'tokens_per_s': clips * 32768 / (ms * 1e-3),
            'model_tflops': clips * flop_clip / (ms * 1e-3) / 1e12, 'model_tflop_per_clip': flop_clip / 1e12,
            'frac_of_bf16_peak': clips * flop_clip / (ms * 1e-3) / 1e12 / tf_peak}


def sampling_benchmark(dev, model, clips, world):
    """Config 5, this rank's shard: `clips` clips x num_steps = 4 new frames (main.py:161 eval_timesteps), each denoised in
    30 mask/replace iterations (one wm_sample_step + one denoiser forward per iteration, captured once), every frame
    decoded by the VQ auto-encoder (codebook gather + conv decoder, 64x64 output).  Clips are independent: the 64-clip
    job is batch-sharded over the ranks with no collective; the caller reduces the time with MAX over ranks."""
    import world_modelz_b200 as wm
    K = C3['num_classes']
    torch.manual_seed(7)
    ae = wm.VqAutoEncoder(64, K, downscale_steps=2, hidden_planes=128, in_channels=1).to(dev).eval()
    tokens = torch.randint(0, K, (clips, *C3['data_shape']), device=dev)
    model.eval()
    num_steps = 4
    wm.sample_frames(model, tokens, num_steps=1, iterations=30, decoder=ae, use_cuda_graph=True)        # capture + warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    frames, decoded = wm.sample_frames(model, tokens, num_steps=num_steps, iterations=30, decoder=ae, use_cuda_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    model.train()
    ok = bool((frames >= 0).all().item() and (frames < K).all().item() and torch.isfinite(decoded[-1]).all().item())
    return ms, {'clips_per_gpu': clips, 'frames_per_clip': num_steps, 'iterations': 30,
                'decoded_frame_shape': list(decoded[-1].shape), 'tokens_ok': ok}


class StdoutGuard:
    """Keeps stdout to the ONE JSON line: while active, file descriptor 1 points at stderr, so banners that native
    libraries write straight to fd 1 (NCCL prints its version there) cannot end up next to the result."""

    def __init__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def release(self):
        if self._saved is not None:
            sys.stdout.flush()
            os.dup2(self._saved, 1)
            os.close(self._saved)
            self._saved = None


# ---------------------------------------------------------------------------------- main
def dbg(msg):
    if os.environ.get('WM_BENCH_DEBUG'):
        print(f'[bench rank {os.environ.get("RANK", "0")} t={time.perf_counter():.1f}] {msg}', file=sys.stderr, flush=True)


def run_b200(args):
    import torch.distributed as dist
    import world_modelz_b200 as wm
    from world_modelz_b200 import ops, parallel
    from world_modelz_b200.denoiser import LossAwareSamplerEma

    guard = StdoutGuard()
    rank, local_rank, world = parallel.init_from_env('nccl')
    dbg(f'process group up: world={world}')
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl b200 needs a CUDA device: there is no CPU fallback')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(42 + rank)
    hbm_gbs, tf_peak, peak_src = peaks()
    B = args.clips_per_gpu
    K = C3['num_classes']

    model = wm.VqVideoDiffusionModel(**C3).to(dev)
    trainer = wm.DenoiserTrainer(model, lr=1e-4, weight_decay=1e-7, compute_dtype=torch.bfloat16,
                                 use_cuda_graph=not args.no_graph)
    dbg('trainer built')
    sampler = LossAwareSamplerEma(seed=42 + rank)
    gen = torch.Generator().manual_seed(1234 + rank)
    n_batches = 4
    host_tokens = [torch.randint(0, K, (B, *C3['data_shape']), generator=gen).pin_memory() for _ in range(n_batches)]
    dev_tokens = [t.to(dev) for t in host_tokens]
    dev_r = [torch.rand(B, device=dev) for _ in range(n_batches)]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync_all()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), t0, t1

    # --- device-resident arm: inputs already in HBM ------------------------------------------
    def step_resident(i):
        trainer.step(dev_tokens[i % n_batches], dev_r[i % n_batches])

    # --- end-to-end arm: pinned host tokens + r in, loss + per-sample losses out, every step ---
    # Two sets of pinned staging buffers: step i's inputs are copied in and its launches enqueued, THEN the host waits
    # for step i-1's losses (already on their way out) and feeds them to the loss-aware sampler -- the reference's
    # sampler also only needs past losses (importance_sampling.py:35-41), so host work overlaps the device step.
    dst_tokens = torch.empty_like(dev_tokens[0])
    dst_r = torch.empty(B, device=dev)
    host_r2 = [torch.empty(B).pin_memory() for _ in range(2)]
    host_loss2 = [torch.empty(1 + B).pin_memory() for _ in range(2)]
    copied2 = [torch.cuda.Event(), torch.cuda.Event()]
    pending = [None]                               # slot of the step whose losses are still in flight

    def drain_e2e():
        if pending[0] is not None:
            j = pending[0]
            copied2[j].synchronize()               # that step's result is on the host
            sampler.update_with_losses(host_r2[j], host_loss2[j][1:])
            pending[0] = None

    def step_e2e(i):
        j = i & 1
        host_r2[j].copy_(sampler.sample(B))
        dst_tokens.copy_(host_tokens[i % n_batches], non_blocking=True)
        dst_r.copy_(host_r2[j], non_blocking=True)
        loss, per_sample = trainer.step(dst_tokens, dst_r)
        host_loss2[j][:1].copy_(loss.reshape(1), non_blocking=True)
        host_loss2[j][1:].copy_(per_sample, non_blocking=True)
        copied2[j].record()
        drain_e2e()                                # previous step's losses -> sampler, while this step runs
        pending[0] = j

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    dbg('warm-up done')
    clocks = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ops.launch_count()
    ms, t0, t1 = timed(step_resident, args.steps)
    launches = args.steps * trainer.launches_per_step() if not args.no_graph else ops.launch_count() - launches0
    clock_info = clocks.stop(t0, t1) if clocks else None
    dbg(f'timed resident arm: {ms:.1f} ms')
    for i in range(3):
        step_e2e(i)
    drain_e2e()

    def run_e2e(i):
        step_e2e(i)
        if i == args.steps - 1:
            drain_e2e()                            # the last step's losses are read inside the timed region too
    ms_e2e, _, _ = timed(run_e2e, args.steps)
    dbg(f'timed e2e arm: {ms_e2e:.1f} ms')

    # --- config 5 on EVERY rank: 8 clips per GPU, batch-sharded, no collective; time = max over ranks ---------------
    samp = samp_pruned = pruned = None
    if not args.no_micro:
        def sampling_arm():
            sync_all()
            samp_ms, info = sampling_benchmark(dev, model, 8, world)
            t_ms = torch.tensor([samp_ms], device=dev)
            if world > 1:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            info.update({'global_clips': 8 * world, 'ms': t_ms.item(),
                         'clips_per_s': 8 * world / (t_ms.item() * 1e-3),
                         'frames_per_s': 8 * world * info['frames_per_clip'] / (t_ms.item() * 1e-3)})
            return info
        samp = sampling_arm()
        # --- opt-in receptive-field cone (NOT the headline): the same step / sampling loop evaluating only the rows the
        # last frame depends on (Local3dAttentionTransformer.last_frame_cone); identical loss and gradients (tests)
        model.prune_receptive_field = True
        samp_pruned = sampling_arm()
        model_p = wm.VqVideoDiffusionModel(**C3).to(dev)
        model_p.prune_receptive_field = True
        trainer_p = wm.DenoiserTrainer(model_p, lr=1e-4, weight_decay=1e-7, compute_dtype=torch.bfloat16,
                                       use_cuda_graph=not args.no_graph)

        def step_pruned(i):
            trainer_p.step(dev_tokens[i % n_batches], dev_r[i % n_batches])
        for i in range(3):
            step_pruned(i)
        ms_p, _, _ = timed(step_pruned, args.steps)
        pruned = {'ms_per_step': ms_p / args.steps, 'clips_per_s': world * B * args.steps / (ms_p * 1e-3),
                  'frames_per_layer': [C3['data_shape'][0] - c for c in model_p.transformer.last_frame_cone(C3['data_shape'][0])[:-1]],
                  'note': 'opt-in (model.prune_receptive_field): layers run on the last frame\'s dependency cone only; '
                          'same loss and parameter gradients as the full step; not used for value / e2e'}
        model.prune_receptive_field = False
        del trainer_p, model_p
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    global_clips = world * B
    value = global_clips * args.steps / (ms * 1e-3)
    e2e_value = global_clips * args.steps / (ms_e2e * 1e-3)
    h2d = B * 4096 * 8 + B * 4
    d2h = (1 + B) * 4
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': config_dict(world, B), 'clocks': clock_info,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'cuda_graph': not args.no_graph, 'peaks': peak_src,
            'train_tokens_per_s': value * 4096,
            'train_model_tflops': value * 42.6e9 / 1e12}
    if not args.no_micro:
        micro = micro_benchmarks(dev, B, hbm_gbs, tf_peak)
        line['roofline'] = micro.pop('roofline')
        # DRAM traffic is not measurable inside an un-profiled run: it comes from the committed `ncu --set full` capture of
        # the same kernels at the same shape, labelled with the commit it was taken at
        traffic_file = os.path.join(ROOT, 'profiles', 'ncu_traffic_r2.json')
        if os.path.exists(traffic_file):
            tr = json.load(open(traffic_file))
            if tr.get('clips') == B:
                line['roofline']['traffic'] = tr['fwd_bwd_dram_bytes']
                line['roofline']['traffic_source'] = tr['source']
                line['roofline']['traffic_commit'] = tr.get('commit')
            if 'vq_dram_bytes' in tr:
                micro['roofline_vq']['traffic'] = tr['vq_dram_bytes']
                micro['roofline_vq']['traffic_commit'] = tr.get('vq_commit', tr.get('commit'))
        line.update(micro)
        line['sampling_config5'] = samp
        line['receptive_field_cone'] = {'train_step': pruned, 'sampling_config5': samp_pruned}
        line['transformer_config4'] = config4_transformer_benchmark(dev, tf_peak)
    else:
        line['roofline'] = None
    if world == 1 and not args.no_cpu_baseline:
        total, n, kind = cpu_train_sample(6, 1, 1)
        cores = os.cpu_count() or 1
        what = 'unmodified reference modules' if kind == 'reference' else f'oracle port ({PORT_VS_REFERENCE})'
        line['cpu_baseline'] = {'value': n / total, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                'sample': f'{n} steps x 1 clip of config 3, fp32, {what} on {cores} host threads'}
    else:
        line['cpu_baseline'] = None
    guard.release()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
