"""N > 1 on real GPUs: world_size-2 NCCL (one process per GPU).  The slab-sharded CUDA model — interface-dof
reduction of the dof-indexed outputs, all-reduced dense kappa blocks, communication-free FE slices of the COO
streams — against the single-GPU CUDA model AND the CPU oracle on the same inputs.  Covers BASELINE configs[4]'s
family (kappa unknowns + 3-D Q1 Poisson control: mixed function / parameter Jacobian blocks,
src/jacobian_functions.jl:135-142, src/GridapPDENLPModel.jl:361-380).  Skipped on a box with fewer than 2 GPUs
(the world-2 gloo tests of test_sharded_gloo.py cover the same host logic on CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-12


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from __graft_entry__ import load_oracle, load_package
    pkg = load_package(); O = load_oracle()
    from pdenlp_b200.model import GridapPDENLPModel
    from pdenlp_b200.sharded import ShardedModel, SlabLayout
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank); dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    res = {}
    try:
        cases = {"kappa3d_u": lambda: pkg.poissonmixed(10, 3, True), "burger_nl": lambda: pkg.burger1d_param(3000),
                 "poissonmixed2": lambda: pkg.poissonmixed2(21), "membrane": lambda: pkg.controlelasticmembrane(40)}
        for name, mk in cases.items():
            spec = mk()
            layout = SlabLayout(spec, world)
            full = GridapPDENLPModel(spec, device=dev)
            loc = GridapPDENLPModel(spec, device=dev, cell_range=layout.ranges[rank])
            sm = ShardedModel(spec, loc, layout, rank)
            orc = O.Oracle(spec, nthreads=2)
            rng = np.random.default_rng(3)
            xh = rng.uniform(-1, 1, full.meta.nvar)
            if spec.nparam:
                xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
            lh = rng.uniform(-1, 1, full.meta.ncon); vh = rng.uniform(-1, 1, full.meta.nvar)
            x, lam, v = (torch.from_numpy(a).to(dev) for a in (xh, lh, vh))

            def rel(a, b):
                a = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
                b = b.cpu().numpy() if isinstance(b, torch.Tensor) else b
                return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.shape == b.shape else float("inf")

            e = {}
            # global COO streams assembled from the per-rank streams (kappa blocks all-reduced inside)
            jl = sm.jac_coord_(x, torch.empty(loc.meta.nnzj, dtype=torch.float64, device=dev))
            hl = sm.hess_coord_(x, lam, torch.empty(loc.meta.nnzh, dtype=torch.float64, device=dev), obj_weight=0.37)
            gj, gh = sm.assemble_global_coo(jl, "jac"), sm.assemble_global_coo(hl, "hess")
            jf, hf = full.jac_coord(x), full.hess_coord(x, lam, obj_weight=0.37)
            e["jac_vs_1gpu"], e["hess_vs_1gpu"] = rel(gj, jf), rel(gh, hf)
            e["jac_vs_oracle"], e["hess_vs_oracle"] = rel(gj, orc.jac_coord(xh)), rel(gh, orc.hess_coord(xh, lh, 0.37))
            k = full.pdemeta.nnzj_k
            e["jac_fe_bitwise"] = 0.0 if torch.equal(gj[k:], jf[k:]) else 1.0
            # the kappa blocks alone (they are tiny next to the FE segments: their own scale)
            pm = full.pdemeta
            if k:
                e["jac_kblock"] = rel(gj[:k], orc.jac_coord(xh)[:k])
                ho = orc.hess_coord(xh, lh, 0.37)
                if pm.nnzh_k_obj:
                    e["hess_kobj"] = rel(gh[: pm.nnzh_k_obj], ho[: pm.nnzh_k_obj])
                a = pm.nnzh_obj
                e["hess_kcon"] = rel(gh[a: a + pm.nnzh_k_con], ho[a: a + pm.nnzh_k_con])
            hl0 = sm.hess_coord_(x, torch.empty(loc.meta.nnzh, dtype=torch.float64, device=dev), obj_weight=1.0)
            e["hess_obj_only"] = rel(sm.assemble_global_coo(hl0, "hess"), orc.hess_coord(xh, None, 1.0))
            z = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
            nv, nc = full.meta.nvar, full.meta.ncon
            e["obj"] = abs(sm.obj(x) - orc.obj(xh)) / max(abs(orc.obj(xh)), 1e-300)
            e["grad"] = rel(sm.allgather_owned(sm.grad_(x, z(nv)), "var"), orc.grad(xh))
            e["cons"] = rel(sm.allgather_owned(sm.cons_(x, z(nc)), "con"), orc.cons(xh))
            e["jprod"] = rel(sm.allgather_owned(sm.jprod_(x, v, z(nc)), "con"), orc.jprod(xh, vh))
            e["jtprod"] = rel(sm.allgather_owned(sm.jtprod_(x, lam, z(nv)), "var"), orc.jtprod(xh, lh))
            e["hprod"] = rel(sm.allgather_owned(sm.hprod_(x, lam, v, z(nv), obj_weight=0.3), "var"), orc.hprod(xh, vh, lh, 0.3))
            res[name] = e
            full.close(); loc.close()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_world2_nccl_sharded_model_matches_single_gpu_and_oracle(cuda_lib, oracle_mod):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    oracle_mod.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs: p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    for rank, res in out:
        for name, errs in res.items():
            bad = {k: v for k, v in errs.items() if not v <= RTOL}
            assert not bad, (rank, name, bad)
