"""Builds libpdenlp_cuda.so in-tree with nvcc for sm_100a (no torch involved:
the library is a plain C-ABI shared object)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libpdenlp_cuda.so")
# pdenlp_cuda.cu holds the C ABI; the ops_*.cu units instantiate disjoint subsets of the (integrand, element)
# kernels and are compiled in parallel
SOURCES = ["pdenlp_cuda.cu", "ops_membrane.cu", "ops_membrane4.cu", "ops_kappa.cu", "ops_misc.cu", "ops_3d.cu", "compaction.cu"]
HEADERS = ["ops.cuh", "kernels.cuh", "integrands.cuh", "jet.cuh", os.path.join("..", "..", "include", "pdenlp_cuda.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        return src, obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = os.path.join(CSRC, "build.log")
    with open(log, "w") as fh:
        for src, obj, cmd, res in results:
            fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    for src, obj, cmd, res in results:
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout[-4000:] + res.stderr[-8000:])
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [r[1] for r in results]
    res = subprocess.run(link, capture_output=True, text=True)
    with open(log, "a") as fh:
        fh.write(" ".join(link) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout[-4000:] + res.stderr[-8000:])
    if verbose:
        print("nvcc ok ->", LIB_PATH, "(ptxas -v log in", log + ")")
    return LIB_PATH
