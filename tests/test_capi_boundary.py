"""The C-ABI library: loads, exports every symbol include/pdenlp_cuda.h declares, the ctypes
mirrors have the C layout, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pdenlp_cuda.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdenlp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(cuda_lib):
    from pdenlp_b200 import capi
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(cuda_lib, s), f"libpdenlp_cuda.so does not export {s}"
    assert sorted(capi.SYMBOLS) == syms
    assert cuda_lib.pdenlp_abi_version() == capi.ABI_VERSION


def test_ctypes_mirrors_match_c_layout(cuda_lib):
    from pdenlp_b200 import capi
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "pdenlp_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pdenlp_desc), offsetof(pdenlp_desc, ncells),
         offsetof(pdenlp_desc, cell_dofs_y), offsetof(pdenlp_desc, wdet), offsetof(pdenlp_desc, coef),
         offsetof(pdenlp_desc, params), sizeof(pdenlp_info), offsetof(pdenlp_desc, nfields_y),
         offsetof(pdenlp_desc, nfields_u), offsetof(pdenlp_desc, field_ndir));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "t.c"); exe = os.path.join(td, "t")
        open(cfile, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    D = capi.Desc
    assert vals == [C.sizeof(D), D.ncells.offset, D.cell_dofs_y.offset, D.wdet.offset, D.coef.offset,
                    D.params.offset, C.sizeof(capi.Info), D.nfields_y.offset, D.nfields_u.offset, D.field_ndir.offset]


def test_header_enum_matches_python_catalogue():
    from pdenlp_b200 import problems as P
    txt = open(HEADER).read()
    enum = dict(re.findall(r"(PDENLP_[A-Z_]+)\s*=\s*(\d+)", txt))
    assert int(enum["PDENLP_MEMBRANE"]) == P.MEMBRANE
    assert int(enum["PDENLP_POISSON_BOLTZMANN"]) == P.POISSON_BOLTZMANN
    assert int(enum["PDENLP_BURGER_LIN"]) == P.BURGER_LIN
    assert int(enum["PDENLP_BURGER_NL"]) == P.BURGER_NL
    assert int(enum["PDENLP_KAPPA_POISSON"]) == P.KAPPA_POISSON


def test_no_cpu_fallback(cuda_lib, pkg):
    """Without a CUDA device pdenlp_create must fail with PDENLP_ECUDA, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pdenlp_b200 import capi, flatten
    fm = flatten.flatten(pkg.controlelasticmembrane(3))
    with pytest.raises(capi.PdenlpError) as ei:
        capi.Handle(fm)
    assert ei.value.code == capi.ECUDA
    from pdenlp_b200.model import GridapPDENLPModel
    with pytest.raises(ValueError):
        GridapPDENLPModel(pkg.controlelasticmembrane(3), device="cpu")


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkgdir = os.path.join(ROOT, "pdenlpmodels.jl_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "load_oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)


def test_sass_has_tma_bulk_store():
    """The coord kernels stream their output with cp.async.bulk (SASS UBLKCP)."""
    from pdenlp_b200 import capi
    try:
        out = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "UBLKCP" in out and "DFMA" in out
