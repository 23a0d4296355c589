"""The C ABI without torch: tests/cabi_client/client.c (plain C, dlopen + cudaMalloc + a user stream, results pre-allocated,
no host sync between the calls -- the way an XLA FFI handler drives the library) is compiled with gcc and
run against the ORACLE-written fixture tests/golden/cabi_fixture.bin (tools/make_cabi_fixture.py)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cabi_client", "client.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(tmp_path):
    exe = os.path.join(str(tmp_path), "cabi_client")
    cmd = ["gcc", "-O2", "-Wall", "-o", exe, SRC, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-ldl", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_c_client_compiles_without_torch(tmp_path):
    """CPU check: the client is plain C against include/tsff.h and the CUDA runtime only."""
    exe = _compile(tmp_path)
    assert os.path.exists(exe)
    src = open(SRC).read()
    assert "torch" not in src.replace("torch-free", "") and "Python.h" not in src


@pytest.mark.gpu
def test_c_client_matches_the_oracle_fixture(tmp_path):
    from tsadar_b200 import _ffi
    exe = _compile(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(CUDA, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, _ffi.LIB_PATH, os.path.join(HERE, "golden", "cabi_fixture.bin")], capture_output=True, text=True,
                       timeout=300, env=env)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout
