"""dist.py on the NCCL backend with the real renderer (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`;
skipped on a single-GPU box).  The CPU twin of the host logic is tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_dist_on_nccl_with_the_real_renderer():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', '29531',
           os.path.join(ROOT, 'tests', 'nccl_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('NCCL_WORKER_OK') == 2, r.stdout[-3000:]
