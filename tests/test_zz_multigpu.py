"""Data-parallel correctness on hardware: world-size-2 NCCL run of the real CUDA model, replicas must stay
bit-identical after N TrainSteps on different per-rank batches.  Needs two GPUs (``gpurun --gpus 2``); skipped on one."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_replicas_stay_bit_identical_over_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_worker.py"), "5"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert "MULTIGPU_OK 2" in proc.stdout
