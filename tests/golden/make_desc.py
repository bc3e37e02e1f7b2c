#!/usr/bin/env python
"""Writes the flattened-model fixtures the plain C test (tests/c/test_sharded.c) reads:

    python tests/golden/make_desc.py      # -> tests/golden/desc_*.bin

The files hold what the host layer hands to pdenlp_create (pdenlp_b200.capi.dump_desc): cell -> dof maps, reference
tables, coefficient samples of three small problems (a C program has no Gridap / Python layer to flatten a model)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    pkg = load_package()
    from pdenlp_b200.capi import dump_desc
    from pdenlp_b200.flatten import flatten
    here = os.path.dirname(os.path.abspath(__file__))
    cases = {"membrane24": pkg.controlelasticmembrane(24), "poissonmixed16": pkg.poissonmixed(16, 2, True),
             "burgerparam512": pkg.burger1d_param(512), "poissonparam16": pkg.poissonparam(16)}
    for name, spec in cases.items():
        path = os.path.join(here, f"desc_{name}.bin")
        dump_desc(flatten(spec), path)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
