"""Alpha-sharded vectors on >= 2 GPUs of one node: parity with the single-GPU engine (runs the torchrun
worker tests/dist_gpu_worker.py).  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_ups_matches_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "dist_gpu_worker.py"),
    ]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(res.stdout[-3000:])
    sys.stderr.write(res.stderr[-3000:])
    assert res.returncode == 0 and "DIST_CHECK_OK" in res.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_sigma_matches_single_gpu(world):
    """H|psi> of an alpha-sharded vector (peer gathers + system-scope atomics over NVLink) against the single-GPU sigma
    kernel, symmetric and unsymmetric integrals; <H> through it against the sharded RDM route."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "dist_sigma_worker.py"),
    ]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(res.stdout[-3000:])
    sys.stderr.write(res.stderr[-3000:])
    assert res.returncode == 0 and "DIST_SIGMA_OK" in res.stdout
