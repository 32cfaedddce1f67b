"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): NCCL all-reduce and the
all-reduce fused into the tail kernel over NVLink peer memory, sharded evaluation and sharded
device-resident optimiser against the unsharded engine (tests/multigpu_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert "MULTIGPU PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
