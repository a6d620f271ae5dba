"""Multi-GPU parity (pytest-collected): launches tests/mgpu_check.py under torchrun on 2, 4 and 8 GPUs when the box has
them (the driver's single-GPU test box skips these; tests/test_chain_gpu.py covers the same machinery there with
several ranks on one device).  mgpu_check compares the NVLink chain with the CPU oracle / a single-GPU run: pruning on
and off, SW and NW, special rows, all chunking modes, repeated calls."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_chain_under_torchrun(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env.setdefault("B200_WATCHDOG_S", "60")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "mgpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(p.stdout[-6000:])
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    assert "MISMATCH" not in p.stdout
