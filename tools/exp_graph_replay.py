import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
pkg = load_package()
from pdenlp_b200.model import GridapPDENLPModel
for n in (724, 2048):
    spec = pkg.controlelasticmembrane(n)
    nlp = GridapPDENLPModel(spec, device="cuda:0", with_coef=False); m = nlp.meta
    x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda(); lam = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, m.ncon)).cuda()
    jv = torch.empty(m.nnzj, dtype=torch.float64, device="cuda"); hv = torch.empty(m.nnzh, dtype=torch.float64, device="cuda")
    def step():
        nlp.jac_coord_(x, jv); nlp.hess_coord_(x, lam, hv, obj_weight=1.0)
    for _ in range(5): step()
    torch.cuda.synchronize()
    K = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): step()
    e1.record(); torch.cuda.synchronize()
    t_plain = e0.elapsed_time(e1) / K
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K): g.replay()
    e1.record(); torch.cuda.synchronize()
    t_graph = e0.elapsed_time(e1) / K
    print(f"n={n}: plain {t_plain:.4f} ms/step, graph replay {t_graph:.4f} ms/step")
    nlp.close(); del jv, hv
