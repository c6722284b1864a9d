"""Multi-GPU check of the data-parallel path on real devices (NCCL): needs >= 2 GPUs (`gpurun --gpus 2`), skipped on a
single-GPU box.  Launches tests/ddp_worker.py under torchrun: overlapped vs plain gradient all-reduce give the same
parameters, replicas stay bit-identical, gradient accumulation (several backward passes per optimizer step) included."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_overlapped_allreduce_matches_plain_and_replicas_stay_identical():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tests", "ddp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("DDP_CHECK") == 8, r.stdout[-3000:]   # 4 scenarios x 2 ranks (lines of the two ranks may interleave)
