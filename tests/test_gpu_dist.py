"""Sharded run == single-GPU batched run, bit for bit, over NCCL on >= 2 GPUs.
Skipped on single-GPU boxes; the host logic is covered on the CPU by tests/test_dist_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_nccl_run_is_bitwise_equal_to_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dist_check ok" in r.stdout


@pytest.mark.gpu
def test_gather_rows_int16_roundtrip_single_process():
    # NCCL has no 16-bit integer type: gather_rows ships int16 as bytes; a world of one must be the identity
    from cmtts_b200.dist import gather_rows
    t = torch.arange(-5, 5, dtype=torch.int16, device="cuda").view(2, 5)
    assert torch.equal(gather_rows(None, t), t)
