#!/usr/bin/env python
"""Measures every BASELINE.json config on one B200 (fills the table of BASELINE.md §3).

For each config: CUDA-event time of every callback named by the config, algorithmic bytes
(SURVEY §8d), achieved GB/s and fraction of the measured HBM peak, and the max relative error of
the CUDA path against the oracle on a slab the oracle finishes in seconds.  Writes one JSON object
per config to stdout (and to --out).  bench.py remains the headline (config 2) line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_oracle, load_package  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]))


def rel(a, b):
    s = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max()) / s


def run(name, spec, callbacks, steps, slab_cells, pk, O):
    from pdenlp_b200.model import GridapPDENLPModel
    t0 = time.perf_counter()
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    setup = time.perf_counter() - t0
    I = nlp.handle.info
    fm = nlp.flat
    nc, nvar, ncon = fm.ncells, I.nvar, I.ncon
    conn = 4 * nc * (fm.ny + fm.nu)
    coefb = 8 * 2 * nc * fm.nq
    rng = np.random.default_rng(1998)
    xh = rng.uniform(-1, 1, nvar)
    if spec.nparam:
        xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
    lh = np.random.default_rng(1999).uniform(-1, 1, ncon)
    vh = np.random.default_rng(2000).uniform(-1, 1, nvar)
    x, lam, v = (torch.from_numpy(a).cuda() for a in (xh, lh, vh))
    jv = torch.empty(I.nnzj, dtype=torch.float64, device="cuda")
    hv = torch.empty(I.nnzh, dtype=torch.float64, device="cuda")
    g = torch.empty(nvar, dtype=torch.float64, device="cuda")
    c = torch.empty(ncon, dtype=torch.float64, device="cuda")
    Hv = torch.empty(nvar, dtype=torch.float64, device="cuda")
    table = {
        "obj": (lambda: nlp.obj(x), 8 * nvar + conn + coefb),
        "grad!": (lambda: nlp.grad_(x, g), 8 * 2 * nvar + conn + coefb),
        "cons!": (lambda: nlp.cons_(x, c), 8 * (nvar + ncon) + conn + coefb),
        "jac_coord!": (lambda: nlp.jac_coord_(x, jv), 8 * I.nnzj + 8 * nvar + conn),
        "hess_coord!(x,y)": (lambda: nlp.hess_coord_(x, lam, hv, obj_weight=1.0), 8 * I.nnzh + 8 * (nvar + ncon) + conn),
        "jprod!": (lambda: nlp.jprod_(x, v, c), 8 * (2 * nvar + ncon) + conn),
        "jtprod!": (lambda: nlp.jtprod_(x, lam, g), 8 * (2 * nvar + ncon) + conn),
        "hprod!(x,y,v)": (lambda: nlp.hprod_(x, lam, v, Hv, obj_weight=1.0), 8 * (3 * nvar + ncon) + conn),
    }
    res = {"config": name, "cells": nc, "nvar": nvar, "ncon": ncon, "nnzj": int(I.nnzj), "nnzh": int(I.nnzh),
           "nq": fm.nq, "nparam": spec.nparam, "setup_s": round(setup, 2), "peak_gbs": pk, "callbacks": {}}
    for cb in callbacks:
        fn, nbytes = table[cb]
        ms = timed(fn, steps)
        res["callbacks"][cb] = {"ms": ms, "cells_per_s": nc / (ms * 1e-3), "algorithmic_bytes": int(nbytes),
                                "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": nbytes / (ms * 1e-3) / 1e9 / pk}
    if "jac_coord!" in callbacks and "hess_coord!(x,y)" in callbacks:
        ms = res["callbacks"]["jac_coord!"]["ms"] + res["callbacks"]["hess_coord!(x,y)"]["ms"]
        res["jac+hess_cells_per_s"] = nc / (ms * 1e-3)
    # parity on the first slab (head of every segment) against the oracle
    sl = min(nc, slab_cells)
    o = O.Oracle(spec, (0, sl))
    pm = nlp.pdemeta
    err = {}
    jo = o.jac_coord(xh)
    nlp.jac_coord_(x, jv)
    jy = jv[pm.nnzj_k: pm.nnzj_k + o.nnzj_y].cpu().numpy()
    err["jac_y_block"] = rel(jy, jo[o.nnzj_k: o.nnzj_k + o.nnzj_y])
    if o.nnzj_u:
        ju = jv[pm.nnzj_k + pm.nnzj_y: pm.nnzj_k + pm.nnzj_y + o.nnzj_u].cpu().numpy()
        err["jac_u_block"] = rel(ju, jo[o.nnzj_k + o.nnzj_y:])
    ho = o.hess_coord(xh, lh, 1.0)
    nlp.hess_coord_(x, lam, hv, obj_weight=1.0)
    a = pm.nnzh_k_obj
    err["hess_obj_fe"] = rel(hv[a: a + o.nnzh_fe].cpu().numpy(), ho[o.nnzh_k_obj: o.nnzh_k_obj + o.nnzh_fe])
    b = pm.nnzh_obj + pm.nnzh_k_con
    ref = ho[o.nnzh_obj + o.nnzh_k_con:]
    got = hv[b: b + o.nnzh_fe].cpu().numpy()
    err["hess_cons_fe"] = float(np.abs(got - ref).max()) / max(float(np.abs(ho).max()), 1e-300)
    res["max_rel_err_vs_oracle_slab"] = err
    res["slab_cells"] = sl
    nlp.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    pkg = load_package()
    O = load_oracle()
    pk = peak()
    coord = ["jac_coord!", "hess_coord!(x,y)"]
    allcb = ["obj", "grad!", "cons!"] + coord + ["jprod!", "jtprod!", "hprod!(x,y,v)"]
    cfgs = [
        ("1: membrane n=100 Q2/Q1", lambda: pkg.controlelasticmembrane(100), allcb, 10000),
        ("2: membrane n=2048 Q2/Q1", lambda: pkg.controlelasticmembrane(2048), allcb, 4096),
        ("3: poissonboltzman2d n=1024 Q1/Q1(L2)", lambda: pkg.poissonboltzman2d(1024), allcb, 8192),
        ("4: burger1d (nonlinear, 1 param) 1M cells", lambda: pkg.burger1d_param(1_000_000), allcb, 65536),
        ("5: kappa + 3D Q1 Poisson control 128^3 (1 GPU)", lambda: pkg.poissonmixed(128, 3, True), allcb, 4096),
    ]
    out = []
    for name, mk, cbs, slab in cfgs:
        if args.only and args.only not in name:
            continue
        r = run(name, mk(), cbs, args.steps, slab, pk, O)
        print(json.dumps(r), flush=True)
        out.append(r)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
