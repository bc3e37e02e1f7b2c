"""Experiment: pure-write bandwidth references vs the coord kernels (config 2)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
from pdenlp_b200.model import GridapPDENLPModel

def timed(fn, steps=20, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        fn(); ev[k + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]))

spec = pkg.controlelasticmembrane(2048)
nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False)
I = nlp.handle.info
x = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, I.nvar)).cuda()
lam = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, I.ncon)).cuda()
jv = torch.empty(I.nnzj, dtype=torch.float64, device="cuda"); hv = torch.empty(I.nnzh, dtype=torch.float64, device="cuda")
gb = lambda n: n * 8 / 1e9
t = timed(lambda: hv.zero_()); print("torch zero_ 6.1GB      %.4f ms  %.0f GB/s" % (t, gb(I.nnzh) / t * 1e3))
t = timed(lambda: hv.fill_(1.5)); print("torch fill_ 6.1GB      %.4f ms  %.0f GB/s" % (t, gb(I.nnzh) / t * 1e3))
t = timed(lambda: hv[: I.nnzh // 2].copy_(hv[I.nnzh // 2: 2 * (I.nnzh // 2)])); print("torch copy 3.05->3.05GB %.4f ms  %.0f GB/s (r+w)" % (t, gb(I.nnzh) / t * 1e3))
t = timed(lambda: nlp.jac_coord_(x, jv)); print("jac_coord              %.4f ms  %.0f GB/s" % (t, gb(I.nnzj) / t * 1e3))
t = timed(lambda: nlp.hess_coord_(x, lam, hv)); print("hess_coord(x,lam)      %.4f ms  %.0f GB/s" % (t, gb(I.nnzh) / t * 1e3))
t = timed(lambda: nlp.hess_coord_(x, hv)); print("hess_coord(x) obj+memset %.4f ms  %.0f GB/s" % (t, gb(I.nnzh) / t * 1e3))
t = timed(lambda: nlp.hess_coord_(x, lam, hv, obj_weight=0.0)); print("hess_coord(x,lam,ow=0) %.4f ms  %.0f GB/s" % (t, gb(I.nnzh) / t * 1e3))
print("nonzeros in cons block:", int((hv[nlp.pdemeta.nnzh_obj:] != 0).sum()), " obj block at ow=0:", int((hv[:nlp.pdemeta.nnzh_obj] != 0).sum()))
