"""Multi-GPU parity (runs only where >= 2 CUDA devices are visible): the real NVLink exchange of the sharded two-site matvec
(fused all-gather epilogue of csrc/gemm.cu + epoch flags of csrc/shard.cu) against the oracle, through torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_matvec_exchange_vs_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + world), os.path.join(ROOT, "tests", "_shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "PARITY OK" in r.stdout, r.stdout[-2000:]
