"""The C ABI from a plain C program (no Python, no torch): tests/cabi_smoke.c is compiled with gcc
against include/fsnerf_b200.h + libfsnerf_b200.so + cudart and run on cuda:0."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(out):
    if not os.path.exists(os.path.join(ROOT, "fsnerf_b200", "libfsnerf_b200.so")):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-O2", "-std=c11", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cabi_smoke.c"), "-o", out, "-L", os.path.join(ROOT, "fsnerf_b200"),
           "-lfsnerf_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + os.path.join(ROOT, "fsnerf_b200"), "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_c_client_compiles_against_the_header(tmp_path):
    """CPU: the header is valid C11 and every symbol the client uses links against the library"""
    r = _compile(str(tmp_path / "cabi_smoke"))
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_c_client_runs(tmp_path):
    exe = str(tmp_path / "cabi_smoke")
    r = _compile(exe)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout)
    assert r.returncode == 0 and "cabi_smoke: OK" in r.stdout, r.stdout + r.stderr
