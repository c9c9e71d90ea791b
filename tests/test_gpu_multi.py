"""Sharded planner on two GPUs == unsharded planner on one (bit identity), through torchrun + NCCL.
Skipped on boxes with a single GPU; tools/check_multi_gpu.py is the program under test."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_planner_is_bit_identical_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTI_GPU_CHECK PASS (world=2)" in res.stdout
