"""Builds liboptimesh_b200.so in-tree with nvcc for sm_100a.

    python -m optimesh_b200.build [--force]

nvcc cross-compiles without a GPU.  Objects are cached by a hash of the sources and
headers, so repeated calls are cheap.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liboptimesh_b200.so")
SOURCES = ["api.cu", "setup.cu", "step.cu", "flip.cu", "stats.cu", "pcg.cu", "loop.cu", "shared.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
    "-diag-suppress", "177,550",
] + os.environ.get("OM_NVCC_EXTRA", "").split()


# step.cu: no implicit contraction -- every instantiation of the step kernels must give a
# vertex the same bits (chain.cuh writes its fused multiply-adds out)
FILE_FLAGS = {"step.cu": ["-fmad=false"]}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _headers_hash() -> str:
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for fn in sorted(os.listdir(d)):
            if fn.endswith((".cuh", ".h")):
                with open(os.path.join(d, fn), "rb") as f:
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, hh: str, force: bool, verbose: bool) -> str:
    path = os.path.join(CSRC, src)
    with open(path, "rb") as f:
        key = hashlib.sha256(f.read() + hh.encode()
                             + " ".join(FILE_FLAGS.get(src, [])).encode()).hexdigest()[:16]
    obj = os.path.join(BUILD, f"{os.path.splitext(src)[0]}.{key}.o")
    if os.path.exists(obj) and not force:
        return obj
    for old in os.listdir(BUILD):
        if old.startswith(os.path.splitext(src)[0] + ".") and old.endswith(".o"):
            os.remove(os.path.join(BUILD, old))
    cmd = [_nvcc(), *NVCC_FLAGS, *FILE_FLAGS.get(src, []), "-c", path, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hh = _headers_hash()
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, hh, force, verbose), SOURCES))
    stamp = os.path.join(BUILD, "link.stamp")
    want = "\n".join(objs)
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read() == want):
        return LIB
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
