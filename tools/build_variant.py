#!/usr/bin/env python
"""Build an experimental variant of libpdenlp_cuda.so with extra -D flags into tools/ab/ (A/B kernel experiments; select
it at run time with PDENLP_CUDA_LIB=<path>).  Usage: python tools/build_variant.py <name> -DFOO=1 [-DBAR=2 ...]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from pdenlp_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "tools", "ab")
obj = os.path.join(out, "obj_" + name)
os.makedirs(obj, exist_ok=True)


def one(src):
    o = os.path.join(obj, os.path.splitext(src)[0] + ".o")
    cmd = [B._nvcc()] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags + ["-c", os.path.join(B.CSRC, src), "-o", o]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr[-4000:])
    return o


with ThreadPoolExecutor(max_workers=len(B.SOURCES)) as pool:
    objs = list(pool.map(one, B.SOURCES))
lib = os.path.join(out, f"libpdenlp_{name}.so")
subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
print(lib)
