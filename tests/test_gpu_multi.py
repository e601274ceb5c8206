"""`-m gpu` tests that need more than one GPU (skipped on a one-GPU box): the column-sharded path of BASELINE config 5
with one process per GPU over NCCL / NVLink peer memory, against the oracle (tests/sharded_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_loops_bit_exact_on_real_gpus(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env.pop("B200LP_GUARD", None)  # symmetric-memory regions are torch's; the guard session flag is not needed here
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "nccl_sharded_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "NCCL sharded parity OK" in p.stdout
