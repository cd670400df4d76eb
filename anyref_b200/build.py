"""In-tree build of libanyref_sam.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m anyref_b200.build [--force] [--verbose]

Every .cu/.cpp under anyref_b200/csrc is compiled to an object (only when stale) and linked into
anyref_b200/libanyref_sam.so.  The built library is git-ignored but travels with `gpurun` snapshots.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libanyref_sam.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libanyref_sam.so")
    return cand


def _stale(src: Path, obj: Path, headers: list[Path]) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *headers])


def _compile(nvcc: str, src: Path, obj: Path, verbose: bool) -> str:
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    log = r.stdout + r.stderr
    (obj.with_suffix(".log")).write_text(log)
    if verbose:
        print(f"--- {src.name}\n{log}")
    return log


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "anyref_sam.h"]
    sources = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))
    jobs = []
    for src in sources:
        obj = OBJ / (src.name + ".o")
        if force or _stale(src, obj, headers):
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs))
    objs = [OBJ / (s.name + ".o") for s in sources]
    if jobs or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
