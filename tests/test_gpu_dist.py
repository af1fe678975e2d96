"""2-rank run of every predictor with distributed=True: NCCL on a box with two GPUs; on a 1-GPU box the two ranks
share cuda:0 and the collectives go through gloo (same sharding / exchange / gather code, host-bounced wire)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_distributed_predictors_match_single_process(tmp_path):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'run_dist_gpu.py'), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert os.path.exists(tmp_path / 'DIST_OK')
