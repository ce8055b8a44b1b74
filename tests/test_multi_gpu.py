"""Launches tests/multi_gpu_check.py under torchrun on 2 GPUs when the box has them (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_sharded_paths(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    with socket.socket() as sk:          # a free port instead of a hard-coded one
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DEIM ok on 2 GPUs" in out.stdout


def test_single_process_multi_gpu_mode(gpu):
    """mor_b200_init_multi: one process, all visible GPUs (also meaningful with one: worker thread, staged ingest)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "single_process_multi_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "single-process multi-GPU check ok" in out.stdout


def test_lazy_multi_gpu_start_from_environment(gpu):
    """MOR_B200_NGPUS: the first call of an uninitialised process starts the multi-GPU mode."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MOR_B200_NGPUS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "single_process_multi_check.py")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "single-process multi-GPU check ok on 2 GPU(s)" in out.stdout
