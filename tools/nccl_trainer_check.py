"""Multi-GPU correctness of the data-parallel trainer on real NCCL (run through torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/nccl_trainer_check.py

Every rank takes its shard of the same global batch for 3 steps (r = 0: no corruption, so the result is
deterministic).  Checks, printed by rank 0 and reflected in the exit code:
  (1) replicas stay BIT-IDENTICAL (master weights, both Adam moments) after the 3 steps;
  (2) the replicas equal a single-GPU trainer stepping on the concatenated batch (fp32 mode: 1e-5; bf16: 2e-2).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import world_modelz_b200 as wm
from world_modelz_b200 import parallel


def main():
    rank, local_rank, world = parallel.init_from_env('nccl')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    kw = dict(data_shape=(4, 8, 8), dim=64, num_classes=32, extents=(1, 2, 2), depth=2, heads=2, dim_head=32, mlp_dim=64)
    per_rank, ok = 4, True
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        torch.manual_seed(100 + rank)                       # DIFFERENT initial weights per rank: rank 0's must win
        model = wm.VqVideoDiffusionModel(**kw).to(dev)
        tr = wm.DenoiserTrainer(model, lr=3e-3, compute_dtype=dtype, use_cuda_graph=True)
        g = torch.Generator().manual_seed(5)
        batches = [torch.randint(0, 32, (world * per_rank, 4, 8, 8), generator=g) for _ in range(3)]
        r = torch.zeros(per_rank, device=dev)
        b, e = parallel.shard_range(world * per_rank, rank, world)
        for full in batches:
            tr.step(full[b:e].to(dev), r)
        torch.cuda.synchronize()
        # (1) bit-identical replicas
        same = True
        for buf in (tr.master, tr.exp_avg, tr.exp_avg_sq):
            ref = buf.clone()
            dist.broadcast(ref, src=0)
            same = same and bool(torch.equal(ref, buf))
        flags = torch.tensor([float(same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        # (2) against one GPU on the concatenated batch (rank 0 only; its initial weights are the replicas')
        err = 0.0
        if rank == 0:
            torch.manual_seed(100)
            solo_model = wm.VqVideoDiffusionModel(**kw).to(dev)
            solo = wm.DenoiserTrainer.__new__(wm.DenoiserTrainer)
            # a trainer that believes it is alone: build it while hiding the process group
            import world_modelz_b200.denoiser as D
            real = D.dist.is_initialized
            D.dist.is_initialized = lambda: False
            try:
                solo.__init__(solo_model, lr=3e-3, compute_dtype=dtype, use_cuda_graph=True)
            finally:
                D.dist.is_initialized = real
            rr = torch.zeros(world * per_rank, device=dev)
            for full in batches:
                solo.step(full.to(dev), rr)
            torch.cuda.synchronize()
            err = ((solo.master - tr.master).abs().max() / solo.master.abs().max()).item()
            print(f'{str(dtype):16s} world={world}: replicas bit-identical={bool(flags.item())}  '
                  f'max|w_dp - w_single|/max|w| = {err:.2e} (bar {tol:.0e})', flush=True)
            ok = ok and bool(flags.item()) and err < tol
        del tr, model
    res = torch.tensor([float(ok)], device=dev)
    dist.broadcast(res, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('NCCL_TRAINER_CHECK', 'PASS' if ok else 'FAIL', flush=True)
    sys.exit(0 if res.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
