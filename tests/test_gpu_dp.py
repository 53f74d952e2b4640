"""Data parallel on real GPUs (needs >= 2): torchrun 2 ranks over NCCL; the sliced-batch run must reproduce the
whole-batch run (global norm statistics, penalty means, gradient all-reduce)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_step_equals_whole_batch_step(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), precision]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "DP_INVARIANCE_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
