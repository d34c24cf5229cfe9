"""Build the CUDA library in-tree:  python -m lanpaint_b200.build [--force] [--verbose]

Plain nvcc, sm_100a only; the resulting `lanpaint_b200/_lib/liblanpaint_b200.so`
is a C-ABI shared library (see include/lanpaint_b200.h) with no torch or Python
dependency, loaded through ctypes by `lanpaint_b200._native`.  Translation units are
compiled in parallel (one nvcc per file) and linked with nvcc -shared.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SOURCES = [os.path.join(CSRC, n) for n in ("lp_substep.cu", "lp_boundary.cu", "lp_misc.cu", "lp_hostnoise.cu", "lp_table.cc")]
HEADERS = [os.path.join(ROOT, "include", "lanpaint_b200.h"), os.path.join(CSRC, "lp_common.cuh")]
LIB_DIR = os.path.join(PKG, "_lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "liblanpaint_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def _newer(path: str, deps) -> bool:
    if not os.path.exists(path):
        return True
    built = os.path.getmtime(path)
    return any(os.path.getmtime(p) > built for p in deps)


def is_stale() -> bool:
    return _newer(LIB_PATH, SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = find_nvcc()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]

    def compile_one(src):
        obj = _obj(src)
        if not force and not _newer(obj, [src] + HEADERS):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", "-o", obj, src]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + res.stdout + res.stderr)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *[o for o, _ in results]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
