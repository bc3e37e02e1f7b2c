import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
pkg = load_package()
from pdenlp_b200.model import GridapPDENLPModel
from tools.bench_configs import timed
for deg, so in ((1, 2), (2, 2), (2, 1)):
    spec = pkg.controlelasticmembrane(2048, degree=deg, state_order=so)
    nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False)
    m = nlp.meta
    x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    lam = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, m.ncon)).cuda()
    jv = torch.empty(m.nnzj, dtype=torch.float64, device="cuda"); hv = torch.empty(m.nnzh, dtype=torch.float64, device="cuda")
    tj = timed(lambda: nlp.jac_coord_(x, jv), 10); th = timed(lambda: nlp.hess_coord_(x, lam, hv, obj_weight=1.0), 10)
    print(f"membrane n=2048 degree={deg} state Q{so}: jac {tj:.3f} ms ({8*m.nnzj/tj/1e6:.0f} GB/s out), hess {th:.3f} ms ({8*m.nnzh/th/1e6:.0f} GB/s out)")
    nlp.close(); del jv, hv
spec = pkg.poissonboltzman2d(1024, degree=2)
nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False); m = nlp.meta
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda(); lam = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, m.ncon)).cuda()
jv = torch.empty(m.nnzj, dtype=torch.float64, device="cuda"); hv = torch.empty(m.nnzh, dtype=torch.float64, device="cuda")
tj = timed(lambda: nlp.jac_coord_(x, jv), 10); th = timed(lambda: nlp.hess_coord_(x, lam, hv, obj_weight=1.0), 10)
print(f"poisson-boltzmann n=1024 degree=2: jac {tj:.3f} ms, hess {th:.3f} ms")
