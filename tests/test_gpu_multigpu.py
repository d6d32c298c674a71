"""SURVEY.md section 8(e) on hardware: the table gathered from N GPUs equals the 1-GPU table byte for byte.
Needs at least two visible GPUs (run it with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multigpu`);
on a one-GPU box it is skipped and the gloo world-2 CPU test (tests/test_host.py) covers the host logic."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gathered_table_equals_single_gpu_table():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    port = 29600 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-GPU table equality ok" in out.stdout
