#!/usr/bin/env python
"""torchrun -n N: the slab-sharded model (NCCL interface-dof reduction) against the single-GPU
model on the same inputs.  Run: python -m torch.distributed.run --nproc-per-node 2 tools/check_sharded_nccl.py"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
from pdenlp_b200.model import GridapPDENLPModel
from pdenlp_b200.sharded import ShardedModel, SlabLayout

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
out = {}
for name, mk in (("membrane64", lambda: pkg.controlelasticmembrane(64)), ("burger_nl", lambda: pkg.burger1d_param(4096)),
                 ("kappa3d", lambda: pkg.poissonmixed(12, 3, True))):
    spec = mk()
    layout = SlabLayout(spec, world)
    full = GridapPDENLPModel(spec, device=dev)
    loc = GridapPDENLPModel(spec, device=dev, cell_range=layout.ranges[rank])
    sm = ShardedModel(spec, loc, layout, rank)
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.uniform(-1, 1, full.meta.nvar)).to(dev)
    lam = torch.from_numpy(rng.uniform(-1, 1, full.meta.ncon)).to(dev)
    v = torch.from_numpy(rng.uniform(-1, 1, full.meta.nvar)).to(dev)
    nv, nc = full.meta.nvar, full.meta.ncon
    def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    errs = {}
    errs["obj"] = abs(sm.obj(x) - full.obj(x)) / abs(full.obj(x))
    z = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
    errs["grad"] = rel(sm.allgather_owned(sm.grad_(x, z(nv)), "var"), full.grad(x))
    errs["cons"] = rel(sm.allgather_owned(sm.cons_(x, z(nc)), "con"), full.cons(x))
    errs["jprod"] = rel(sm.allgather_owned(sm.jprod_(x, v, z(nc)), "con"), full.jprod(x, v))
    errs["jtprod"] = rel(sm.allgather_owned(sm.jtprod_(x, lam, z(nv)), "var"), full.jtprod(x, lam))
    errs["hprod"] = rel(sm.allgather_owned(sm.hprod_(x, lam, v, z(nv), obj_weight=0.3), "var"), full.hprod(x, lam, v, obj_weight=0.3))
    # COO slices: gather sizes, compare this rank's slice with the full stream
    off = sm.global_segment_offsets()
    pm, fpm = loc.pdemeta, full.pdemeta
    jl, jf = loc.jac_coord(x), full.jac_coord(x)
    a = fpm.nnzj_k + off["jy"]
    ok = torch.equal(jl[pm.nnzj_k: pm.nnzj_k + pm.nnzj_y], jf[a: a + pm.nnzj_y])
    b = fpm.nnzj_k + fpm.nnzj_y + off["ju"]
    ok &= torch.equal(jl[pm.nnzj_k + pm.nnzj_y:], jf[b: b + pm.nnzj_u])
    hl, hf = loc.hess_coord(x, lam, obj_weight=0.3), full.hess_coord(x, lam, obj_weight=0.3)
    a = fpm.nnzh_k_obj + off["hfe"]
    ok &= torch.equal(hl[pm.nnzh_k_obj: pm.nnzh_k_obj + pm.nnzh_fe], hf[a: a + pm.nnzh_fe])
    b = fpm.nnzh_obj + fpm.nnzh_k_con + off["hfe"]
    ok &= torch.equal(hl[pm.nnzh_obj + pm.nnzh_k_con:], hf[b: b + pm.nnzh_fe])
    errs["coo_slices_bitwise"] = bool(ok)
    errs["iface_dofs"] = int(layout.iface_var.size)
    out[name] = errs
    full.close(); loc.close()
res = [None] * world
dist.all_gather_object(res, out)
if rank == 0:
    bad = 0
    for r, o in enumerate(res):
        for name, e in o.items():
            print(f"rank {r} {name}: " + ", ".join(f"{k}={v:.1e}" if isinstance(v, float) else f"{k}={v}" for k, v in e.items()))
            bad += any(isinstance(v, float) and v > 1e-12 for v in e.values()) or not e["coo_slices_bitwise"]
    print("SHARDED_NCCL_OK" if not bad else "SHARDED_NCCL_FAIL")
dist.destroy_process_group()
