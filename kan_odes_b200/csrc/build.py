"""Builds libkanode_b200.so in-tree with nvcc for sm_100a (no torch types, plain C ABI)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "libkanode_b200.so"
SOURCES = ["kanode_api.cu"]
DEPS = ["kanode_api.cu", "kanode_host.h", "kanode_math.cuh", "kanode_small.cuh", "kanode_generic.cuh", "kanode_small_ls.cuh", "kanode_wide.cuh",
        "../../include/kanode.h"]
NVCC_FLAGS = ["-std=c++20", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any((HERE / d).stat().st_mtime > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(OUT), *SOURCES]
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    (HERE / "build.log").write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libkanode_b200.so")
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
