"""Generates tests/golden/*.npz.

PROVENANCE: the Julia reference (PDENLPModels.jl on Gridap 0.15.5) cannot run in this image, so
these vectors come from the CPU oracle (oracle/pdenlp_oracle.cpp), NOT from the reference itself.
They freeze the oracle's output (regression pin for both the oracle and the CUDA path); true
reference vectors can replace them out of band with julia/export_golden.jl, which writes the
same keys.  Run:  python tests/golden/make_golden.py [case names; default = all]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from __graft_entry__ import load_oracle, load_package  # noqa: E402

CASES = {
    "membrane_n3": lambda p: p.controlelasticmembrane(3),
    "poissonboltzman_l2_n4": lambda p: p.poissonboltzman2d(4),
    "burger1d_n6": lambda p: p.burger1d(6),
    "burger1d_param_n6": lambda p: p.burger1d_param(6),
    "poissonmixed_n3": lambda p: p.poissonmixed(3),
    "inversepoisson3d_n2": lambda p: p.poissonmixed(2, 3, True),
    "penalizedpoisson_n4": lambda p: p.penalizedpoisson(4),
    "basicunconstrained_n3": lambda p: p.basicunconstrained(3),
    # multi-field spaces
    "controlsir_n6": lambda p: p.controlsir(6),
    "cellincrease_n5": lambda p: p.cellincrease(5),
    "dynamicsir_n6": lambda p: p.dynamicsir(6),
    "torebrachistochrone_n4": lambda p: p.torebrachistochrone(4),
    "poissonmixed2_n3": lambda p: p.poissonmixed2(3),
    # NoFETerm objectives
    "poissonparam_n3": lambda p: p.poissonparam(3),
    "sis_n6": lambda p: p.sis(6),
}


def inputs(nvar, ncon, nparam):
    x = np.random.default_rng(1998).uniform(-1, 1, nvar)
    if nparam:
        x[:nparam] = 1 + 0.1 * np.random.default_rng(1997).uniform(-1, 1, nparam)
    lam = np.random.default_rng(1999).uniform(-1, 1, ncon)
    v = np.random.default_rng(2000).uniform(-1, 1, nvar)
    return x, lam, v


def main():
    pkg = load_package(); O = load_oracle()
    only = sys.argv[1:]
    for name, mk in CASES.items():
        if only and name not in only:
            continue
        spec = mk(pkg)
        o = O.Oracle(spec, nthreads=1)
        x, lam, v = inputs(o.nvar, o.ncon, spec.nparam)
        jr, jc = o.jac_structure(); hr, hc = o.hess_structure()
        if spec.unconstrained:  # objective only: ncon = 0, no Jacobian, Hessian = objective segment
            no = o.nnzh_obj
            np.savez_compressed(
                os.path.join(HERE, name + ".npz"),
                x=x, v=v, obj_weight=0.37, obj=o.obj(x), grad=o.grad(x), unconstrained=True,
                hess_rows=hr[:no], hess_cols=hc[:no], hess_obj_vals=o.hess_coord(x, None, 0.37)[:no],
                hprod=o.hprod(x, v, None, 0.37), nnzh_obj=no,
            )
            print(name, "nvar", o.nvar, "unconstrained, nnzh", no)
            continue
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            x=x, lam=lam, v=v, obj_weight=0.37, obj=o.obj(x), grad=o.grad(x), cons=o.cons(x),
            jac_rows=jr, jac_cols=jc, jac_vals=o.jac_coord(x),
            hess_rows=hr, hess_cols=hc, hess_vals=o.hess_coord(x, lam, 0.37), hess_obj_vals=o.hess_coord(x, None, 0.37),
            jprod=o.jprod(x, v), jtprod=o.jtprod(x, lam), hprod=o.hprod(x, v, lam, 0.37), nnzh_obj=o.nnzh_obj,
        )
        print(name, "nvar", o.nvar, "nnzj", o.nnzj, "nnzh", o.nnzh)


if __name__ == "__main__":
    main()
