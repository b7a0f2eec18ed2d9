"""Multi-GPU gradient equality on the CUDA path (NCCL): the all-reduced gradient of the ray-sharded
step equals the single-GPU gradient of the same global batch.  Needs >= 2 devices
on the box; skipped otherwise (the host-side sharding logic is covered on CPU with gloo in
tests/test_parallel_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 4, 8])
def test_data_parallel_gradient_equals_single_gpu(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + n), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert f"dp{n} vs single GPU" in r.stdout
