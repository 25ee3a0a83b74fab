"""Two-rank NCCL path on real GPUs (skipped on boxes with fewer than two): tools/multi_gpu_check.py under torchrun --
the stellar phase and the complete panchromatic flow (stellar -> self-absorption cycles -> dust emission)."""
import os
import subprocess
import sys

import pytest

import common


@pytest.mark.gpu
def test_two_ranks_allreduce_matches_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(common.ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU check ok" in r.stdout
    # stellar -> self-absorption cycles -> dust emission on two ranks equals the one-rank run of the same Philox streams
    assert "multi-GPU pan flow ok on 2 ranks (3 fixed cycles)" in r.stdout
    assert "multi-GPU pan flow ok on 2 ranks (cycles until convergence)" in r.stdout
