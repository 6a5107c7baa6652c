"""Coupled (summed-density) mode across ranks: 2 processes, one GPU each, the density all-reduced by NCCL inside
libmsm_b200, every rank's streams checked against the CPU oracle ensemble (tests/summed_multi_gpu_worker.py).
Skipped on boxes with a single GPU; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(gpu_count() < 2, reason="needs 2 GPUs")
def test_summed_mode_on_two_ranks_matches_the_oracle_ensemble():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "summed_multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("-> OK") >= 14 and "FAIL" not in out.stdout
