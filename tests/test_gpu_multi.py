"""GPU tier, >= 2 GPUs on the box (skipped otherwise): the sampler's collectives over NCCL - tools/nccl_checks.py under
torchrun with world size 2 (dopri5 with the shared error norm == the unsharded run, statistics all-reduce, sample gather,
global IQR mask)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_nccl_collectives_world_size_2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(REPO, "tools", "nccl_checks.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    print(res.stdout[-3000:], res.stderr[-2000:])
    assert res.returncode == 0 and "NCCL CHECKS PASSED" in res.stdout
