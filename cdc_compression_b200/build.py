"""Compile csrc/ into libcdc_b200.so for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/engine.cu"]
HEADERS = ["csrc/common.cuh", "csrc/igemm_hmma.cuh", "csrc/igemm_tc.cuh", "csrc/attn.cuh", "csrc/attn_tc.cuh", "csrc/final_tc.cuh", "csrc/misc.cuh", "../include/cdc_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(_HERE, f)) > t for f in SOURCES + HEADERS
               if os.path.exists(os.path.join(_HERE, f)))


def build(force: bool = False, verbose: bool = False) -> str:
    out = os.path.join(_HERE, "libcdc_b200.so")
    if not force and not _stale(out):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *[os.path.join(_HERE, s) for s in SOURCES], "-o", out]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("CDC_ROLE_CLK"):
        cmd.insert(1, "-DCDC_ROLE_CLK")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
