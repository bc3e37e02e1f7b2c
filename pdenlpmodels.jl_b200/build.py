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
HEADERS = ["ops.cuh", "kernels.cuh", "kvector.cuh", "kmma.cuh", "integrands.cuh", "jet.cuh", os.path.join("..", "..", "include", "pdenlp_cuda.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def source_hash() -> str:
    """Build id: sha256 over every source and header of the library (names + contents, fixed order).  It is
    baked into the binary at compile time (-DPDENLP_BUILD_ID, returned by pdenlp_build_id()); capi.load()
    recomputes it from the tree and refuses a library built from other sources."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES) + sorted(HEADERS):
        path = os.path.normpath(os.path.join(CSRC, f))
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def built_id(path: str = None) -> str:
    """The build id baked into an existing library ('' if none).  Read from the file's bytes (marker string
    "PDENLP_BUILD_ID=<id>;"), not through dlopen: a library already mapped in this process would answer for
    the file it was loaded from, not for the file on disk now."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        return ""
    import mmap
    marker = b"PDENLP_BUILD_ID="
    with open(path, "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
        i = mm.find(marker)
        if i < 0:
            return ""
        j = mm.find(b";", i)
        return mm[i + len(marker): j].decode(errors="replace") if 0 < j - i < 64 else ""


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    """True when the library is missing or was built from other sources (content hash, not mtimes)."""
    return built_id() != source_hash()


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    bid = source_hash()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        extra = [f'-DPDENLP_BUILD_ID="{bid}"'] if src == "pdenlp_cuda.cu" else []
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
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
    if built_id() != bid:
        raise RuntimeError("the freshly built library does not report the build id of the sources")
    if verbose:
        print("nvcc ok ->", LIB_PATH, "build id", bid, "(ptxas -v log in", log + ")")
    return LIB_PATH
