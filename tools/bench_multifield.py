#!/usr/bin/env python
"""Times the multi-field (ODE) problems at 10^6 cells on one GPU: every callback, CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from tools.bench_configs import timed  # noqa: E402


def main():
    pkg = load_package()
    from pdenlp_b200.model import GridapPDENLPModel
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    out = []
    for name in ("controlsir", "sis", "cellincrease", "dynamicsir", "torebrachistochrone"):
        spec = getattr(pkg, name)(n)
        nlp = GridapPDENLPModel(spec, device="cuda:0")
        m = nlp.meta
        rng = np.random.default_rng(0)
        x = torch.from_numpy(rng.uniform(0.1, 1, m.nvar)).cuda()
        lam = torch.from_numpy(rng.uniform(-1, 1, m.ncon)).cuda()
        v = torch.from_numpy(rng.uniform(-1, 1, m.nvar)).cuda()
        hv = torch.empty(m.nnzh, dtype=torch.float64, device="cuda")
        g = torch.empty(m.nvar, dtype=torch.float64, device="cuda")
        Hv = torch.empty(m.nvar, dtype=torch.float64, device="cuda")
        res = {"problem": name, "cells": n, "nvar": m.nvar, "ncon": m.ncon, "nnzj": m.nnzj, "nnzh": m.nnzh}
        res["grad!"] = timed(lambda: nlp.grad_(x, g), 10)
        if m.ncon:
            jv = torch.empty(m.nnzj, dtype=torch.float64, device="cuda")
            c = torch.empty(m.ncon, dtype=torch.float64, device="cuda")
            res["cons!"] = timed(lambda: nlp.cons_(x, c), 10)
            res["jac_coord!"] = timed(lambda: nlp.jac_coord_(x, jv), 10)
            res["hess_coord!(x,y)"] = timed(lambda: nlp.hess_coord_(x, lam, hv, obj_weight=1.0), 10)
            res["jprod!"] = timed(lambda: nlp.jprod_(x, v, c), 10)
            res["jtprod!"] = timed(lambda: nlp.jtprod_(x, lam, g), 10)
            res["hprod!(x,y,v)"] = timed(lambda: nlp.hprod_(x, lam, v, Hv, obj_weight=1.0), 10)
            res["jac_GBs"] = 8 * m.nnzj / res["jac_coord!"] / 1e6
            res["hess_GBs"] = 8 * m.nnzh / res["hess_coord!(x,y)"] / 1e6
        else:
            res["hess_coord!(x)"] = timed(lambda: nlp.hess_coord_(x, hv, obj_weight=1.0), 10)
            res["hprod!(x,v)"] = timed(lambda: nlp.hprod_(x, v, Hv, obj_weight=1.0), 10)
        print(json.dumps(res), flush=True)
        out.append(res)
        nlp.close()
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
