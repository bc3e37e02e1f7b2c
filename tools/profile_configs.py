#!/usr/bin/env python
"""One launch of every callback of one BASELINE config, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'k_(vector|jac|hess|zero)' -c 40 \
        -o gpurun_out/r02_cfg5 python tools/profile_configs.py --config 5

(tools/ncu_configs_summary.py condenses the reports into profiles/r02_configs_ncu_summary.json, the file
bench.py reads the FP64-pipe utilisation from.)"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="5")
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    pkg = load_package()
    from pdenlp_b200.model import GridapPDENLPModel
    spec = (bench.OTHER_CONFIGS[args.config]["make"](pkg) if args.config != "2" else pkg.controlelasticmembrane(2048))
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    I = nlp.handle.info
    rng = np.random.default_rng(1998)
    xh = rng.uniform(-1, 1, I.nvar)
    if spec.nparam:
        xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
    x = torch.from_numpy(xh).cuda()
    lam = torch.from_numpy(np.random.default_rng(1999).uniform(-1, 1, I.ncon)).cuda()
    v = torch.from_numpy(np.random.default_rng(2000).uniform(-1, 1, I.nvar)).cuda()
    new = lambda n: torch.empty(int(n), dtype=torch.float64, device="cuda")
    jv, hv, g, c, Hv = new(I.nnzj), new(I.nnzh), new(I.nvar), new(I.ncon), new(I.nvar)
    for _ in range(args.reps):
        nlp.obj(x); nlp.grad_(x, g); nlp.cons_(x, c)
        nlp.jac_coord_(x, jv); nlp.hess_coord_(x, lam, hv, obj_weight=1.0)
        nlp.jprod_(x, v, c); nlp.jtprod_(x, lam, g); nlp.hprod_(x, lam, v, Hv, obj_weight=1.0)
    torch.cuda.synchronize()
    print("profile_configs ok", args.config)


if __name__ == "__main__":
    main()
