"""The C ABI is usable without Python or torch: tests/c_abi/abi_smoke.cu is a plain CUDA-runtime program
that links liblanpaint_b200.so, drives a whole outer step through raw pointers and checks it against a
scalar restatement.  CPU: it must compile, link and load (`--link-only`).  GPU: it must agree."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT
from lanpaint_b200 import _native

SRC = os.path.join(ROOT, "tests", "c_abi", "abi_smoke.cu")
LIBDIR = os.path.dirname(_native.lib_path())


def _build(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    _native.load()
    exe = str(tmp_path / "abi_smoke")
    cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"),
           "-o", exe, SRC, "-L", LIBDIR, "-llanpaint_b200", "-Xlinker", f"-rpath={LIBDIR}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return exe


def test_c_consumer_compiles_links_and_loads(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe, "--link-only"], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip() == f"abi {_native.ABI_VERSION}", res.stdout + res.stderr


@pytest.mark.gpu
def test_c_consumer_matches_closed_form(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "abi_smoke: max abs err" in res.stdout
