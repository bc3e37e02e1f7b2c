#!/usr/bin/env python
"""One-off 64-bit indexing check: membrane n = 4096 (16.8 M cells, nnzh = 3.05e9 > 2^31 entries, 24 GB of Hessian
values).  Compares the LAST cells' runs (offsets beyond 2^31) with the oracle and looks for unwritten entries."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_oracle, load_package  # noqa: E402


def main():
    pkg = load_package(); O = load_oracle()
    from pdenlp_b200.model import GridapPDENLPModel
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    t0 = time.perf_counter()
    spec = pkg.controlelasticmembrane(n)
    nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False)
    m = nlp.meta
    print(f"n={n}: cells={spec.model.ncells} nvar={m.nvar} nnzj={m.nnzj} nnzh={m.nnzh} (2^31={2**31}) setup {time.perf_counter()-t0:.1f}s", flush=True)
    assert m.nnzh > 2 ** 31
    rng = np.random.default_rng(1998)
    xh = rng.uniform(-1, 1, m.nvar); lh = rng.uniform(-1, 1, m.ncon)
    x, lam = torch.from_numpy(xh).cuda(), torch.from_numpy(lh).cuda()
    hv = torch.full((m.nnzh,), float("nan"), dtype=torch.float64, device="cuda")
    nlp.hess_coord_(x, lam, hv, obj_weight=0.37)
    torch.cuda.synchronize()
    # no entry left unwritten (chunked reduction: isnan over 24 GB)
    nan = 0
    for a in range(0, m.nnzh, 1 << 28):
        nan += int(torch.isnan(hv[a: a + (1 << 28)]).sum())
    assert nan == 0, nan
    nc = spec.model.ncells
    c0 = nc - 8192
    o = O.Oracle(spec, (c0, nc))
    ho = o.hess_coord(xh, lh, 0.37)
    pm = nlp.pdemeta
    a = pm.nnzh_fe - o.nnzh_fe              # the slab's run is the tail of the objective FE segment
    assert a > 2 ** 31 - 10 ** 9
    got = hv[a: a + o.nnzh_fe].cpu().numpy()
    err_o = np.abs(got - ho[: o.nnzh_fe]).max() / np.abs(ho[: o.nnzh_fe]).max()
    b = pm.nnzh_obj + pm.nnzh_fe - o.nnzh_fe   # tail of the constraint FE segment: offsets beyond 2^31
    assert b > 2 ** 31
    got = hv[b: b + o.nnzh_fe].cpu().numpy()
    err_c = np.abs(got - ho[o.nnzh_fe:]).max()
    print(f"tail slab vs oracle: obj block rel err {err_o:.2e} (offset {a}), cons block abs err {err_c:.2e} (offset {b})")
    assert err_o <= 1e-12 and err_c == 0.0
    del hv
    jv = torch.full((m.nnzj,), float("nan"), dtype=torch.float64, device="cuda")
    nlp.jac_coord_(x, jv)
    torch.cuda.synchronize()
    jo = o.jac_coord(xh)
    a = pm.nnzj_y - o.nnzj_y
    e1 = np.abs(jv[a: a + o.nnzj_y].cpu().numpy() - jo[: o.nnzj_y]).max() / np.abs(jo[: o.nnzj_y]).max()
    b = pm.nnzj_y + pm.nnzj_u - o.nnzj_u
    e2 = np.abs(jv[b: b + o.nnzj_u].cpu().numpy() - jo[o.nnzj_y:]).max() / np.abs(jo[o.nnzj_y:]).max()
    print(f"jac tail slab vs oracle: y-block {e1:.2e}, u-block {e2:.2e}")
    assert e1 <= 1e-12 and e2 <= 1e-12 and not bool(torch.isnan(jv[:: 1009]).any())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hv = torch.empty(m.nnzh, dtype=torch.float64, device="cuda")
    nlp.hess_coord_(x, lam, hv, obj_weight=1.0); nlp.jac_coord_(x, jv)
    ev0.record()
    for _ in range(5):
        nlp.jac_coord_(x, jv); nlp.hess_coord_(x, lam, hv, obj_weight=1.0)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 5
    print(f"step {ms:.3f} ms = {nc / ms / 1e6:.3f}e9 cells/s")
    print("LARGE_OK")


if __name__ == "__main__":
    main()
