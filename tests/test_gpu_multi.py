"""Two-rank NCCL run of the data-parallel trainer (needs >= 2 GPUs; skipped on a one-GPU box).  The same script is
run through `gpurun --gpus 2` and its log kept in profiles/nccl_trainer_check_r2.log."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_nccl_trainer_replicas_identical_and_match_single_gpu():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tools', 'nccl_trainer_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'NCCL_TRAINER_CHECK PASS' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
