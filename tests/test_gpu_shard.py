"""Multi-GPU (>= 2 devices on one box) check of the intra-sample sharding path over NCCL."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_matches_single_gpu():
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "shard_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARD_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
