"""Multi-GPU parity (skipped on boxes with fewer GPUs): NCCL halo exchange + fused kernels
against the serial oracle, one torchrun rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shared", ["0", "1", "masked"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_jacobian_matches_oracle(world, shared):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1", "--nnodes=1",
           f"--nproc-per-node={world}", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MGPU_SHARED=shared))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_halo_matches_oracle(world):
    """sum-and-share written straight into the neighbours' windows over NVLink (csrc/b200_halo.cu)"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1", "--nnodes=1",
           f"--nproc-per-node={world}", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MGPU_SHARED="masked", MGPU_HALO="p2p"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout and "halo p2p" in r.stdout and "bit-identical: True" in r.stdout
