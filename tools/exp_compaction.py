"""Full-size timing of the COO -> CSC compaction (config 2) and of the affine linear API."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
from pdenlp_b200.model import GridapPDENLPModel

def timed(fn, steps=10, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        fn(); ev[k + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]))

spec = pkg.controlelasticmembrane(2048, affine=True)
t0 = time.perf_counter()
nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=True)
torch.cuda.synchronize()
print("affine model create incl. CSC plan over %d COO entries: %.2f s; nnz(A) = %d" % (nlp.handle.info.nnzj, time.perf_counter() - t0, nlp.meta.nnzj))
I = nlp.handle.info
x = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, I.nvar)).cuda()
v = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, I.nvar)).cuda()
w = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, I.ncon)).cuda()
vals = torch.empty(nlp.meta.nnzj, dtype=torch.float64, device="cuda")
c = torch.empty(I.ncon, dtype=torch.float64, device="cuda"); g = torch.empty(I.nvar, dtype=torch.float64, device="cuda")
print("jac_lin_coord! (copy of cached findnz values) %.4f ms" % timed(lambda: nlp.jac_coord_(x, vals)))
print("cons_lin!  %.4f ms   jprod_lin! %.4f ms   jtprod_lin! %.4f ms" % (timed(lambda: nlp.cons_(x, c)), timed(lambda: nlp.jprod_(x, v, c)), timed(lambda: nlp.jtprod_(x, w, g))))
coo = torch.empty(I.nnzj, dtype=torch.float64, device="cuda")
nlp.handle.jac_coord(x.data_ptr(), coo.data_ptr())
t = timed(lambda: nlp._jac_plan.apply(coo.data_ptr(), vals.data_ptr(), torch.cuda.current_stream().cuda_stream))
print("compaction apply %d -> %d entries: %.4f ms  (%.0f GB/s on 16 B/COO entry + 8 B/out entry)" % (I.nnzj, nlp.meta.nnzj, t, (I.nnzj * 16 + nlp.meta.nnzj * 8) / t / 1e6))
h = torch.empty(nlp.meta.nnzh, dtype=torch.float64, device="cuda")
print("hess_coord! (objective part only, %d entries) %.4f ms" % (nlp.meta.nnzh, timed(lambda: nlp.hess_coord_(x, w, h))))
