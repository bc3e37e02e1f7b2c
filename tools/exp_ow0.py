import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
from pdenlp_b200.model import GridapPDENLPModel
spec = pkg.controlelasticmembrane(2048)
nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False)
I = nlp.handle.info
x = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, I.nvar)).cuda()
lam = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, I.ncon)).cuda()
hv = torch.empty(I.nnzh, dtype=torch.float64, device="cuda")
for _ in range(4):
    nlp.hess_coord_(x, lam, hv, obj_weight=0.0)
torch.cuda.synchronize()
