"""Multi-GPU parity tests: the point, toy and event splits of blueice_b200.distributed on real
GPUs over NCCL + NVLink peer memory, bit for bit against single-GPU evaluations of the same inputs.  One process per
GPU via torchrun (tests/multi_gpu_worker.py); skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_evaluations_match_one_gpu(world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stdout[-3000:] + "\n" + res.stderr[-6000:]
    assert "MULTI_GPU_OK" in res.stdout, res.stdout[-3000:]
