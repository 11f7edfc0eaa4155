"""Builds libkanode_b200.so in-tree with nvcc for sm_100a (no torch types, plain C ABI).

Each translation unit is compiled to its own object (in parallel, rebuilt only when one of its dependencies changed) and
the objects are linked into the shared library."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "libkanode_b200.so"
BUILD = HERE / "_build"
COMMON = ["kanode_host.h", "kanode_math.cuh", "kanode_wide_api.h", "../../include/kanode.h"]
UNITS = {
    "kanode_api.cu": ["kanode_small.cuh", "kanode_small_host.h", "kanode_generic.cuh"],
    "kanode_lg.cu": ["kanode_small.cuh", "kanode_small_host.h", "kanode_small_lg.cuh"],
    "kanode_wide.cu": ["kanode_wide.cuh", "kanode_wsrc.cuh"],
    "kanode_peer.cu": [],
}
SOURCES = list(UNITS)
NVCC_FLAGS = ["-std=c++20", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _obj(src: str) -> Path:
    return BUILD / (Path(src).stem + ".o")


def _log(src: str) -> Path:
    return BUILD / (Path(src).stem + ".log")


def _stale(src: str) -> bool:
    o = _obj(src)
    if not o.exists():
        return True
    t = o.stat().st_mtime
    return any((HERE / d).stat().st_mtime > t for d in [src, *UNITS[src], *COMMON])


def needs_build() -> bool:
    if not OUT.exists() or any(_stale(s) for s in SOURCES):
        return True
    return any(_obj(s).stat().st_mtime > OUT.stat().st_mtime for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    BUILD.mkdir(exist_ok=True)
    procs = []
    for s in SOURCES:
        if force or _stale(s):
            cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(_obj(s)), s]
            procs.append((s, subprocess.Popen(cmd, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for s, p in procs:
        out, _ = p.communicate()
        _log(s).write_text(out)
        if p.returncode != 0:
            failed.append(s)
            sys.stderr.write(out)
    (HERE / "build.log").write_text("\n".join(f"==== {s} ====\n{_log(s).read_text()}" for s in SOURCES if _log(s).exists()))
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    r = subprocess.run([nvcc, "-shared", "-o", str(OUT), *[str(_obj(s)) for s in SOURCES]], cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking libkanode_b200.so failed")
    if verbose:
        print((HERE / "build.log").read_text())
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
