"""The forward kernel launched as clusters of two CTAs that share multicast weight stages
(FSNERF_FWD_PAIR=1, csrc/mlp_fwd2.cu): same parity bars as the default launch.  The switch is
read once per process, so the MLP parity tests are re-run in a child process with it set; their
sizes cover odd tile counts, where the odd CTA of the last pair computes a dummy tile."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_forward_cta_pairs_pass_the_mlp_parity_tests():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, FSNERF_FWD_PAIR="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_kernels.py", "tests/test_gpu_train.py", "-x", "-q",
                        "-m", "gpu", "-k", "mlp_forward or mlp_backward or train_step or render or drop"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1000:])
    assert " passed" in r.stdout
