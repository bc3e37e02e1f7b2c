"""Multi-GPU behind the C ABI (pdenlp_sharded_*, include/pdenlp_cuda.h): a plain C program (tests/c/test_sharded.c) drives
the devices through the header only — no Python at run time — and must reproduce the single-device results."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pdenlpmodels.jl_b200", "csrc")
FIXTURES = ["membrane24", "poissonmixed16", "burgerparam512", "poissonparam16"]


def _build(tmp_path):
    exe = str(tmp_path / "test_sharded")
    cmd = ["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "test_sharded.c"),
           "-L", CSRC, "-lpdenlp_cuda", "-lm", f"-Wl,-rpath,{CSRC}", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_driver_compiles_and_links_against_the_header_only(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    for f in FIXTURES:
        assert os.path.getsize(os.path.join(ROOT, "tests", "golden", f"desc_{f}.bin")) > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", FIXTURES)
def test_c_driver_sharded_equals_single_device(tmp_path, fixture):
    import torch
    exe = _build(tmp_path)
    ngpu = torch.cuda.device_count()
    for ndev in sorted({1, min(2, ngpu), min(8, ngpu)}):
        res = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", f"desc_{fixture}.bin"), str(ndev)],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0 and f"SHARDED_C_OK ndev={ndev}" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


@pytest.mark.gpu
def test_sharded_handle_matches_the_oracle(pkg):
    """The same entry points through ctypes on a problem with kappa blocks: global COO streams, structures and the
    dof-indexed outputs against the CPU oracle over the whole mesh."""
    import torch
    from __graft_entry__ import load_oracle
    from pdenlp_b200.capi import ShardedHandle
    from pdenlp_b200.flatten import flatten
    O = load_oracle()
    spec = pkg.poissonmixed(12, 2, True)
    orc = O.Oracle(spec)
    g = np.random.default_rng(11)
    x = g.uniform(-1, 1, orc.nvar); x[:2] = 1 + 0.1 * g.uniform(-1, 1, 2)
    lam, v = g.uniform(-1, 1, orc.ncon), g.uniform(-1, 1, orc.nvar)
    ndev = min(2, torch.cuda.device_count())
    h = ShardedHandle(flatten(spec), ndev)
    assert (h.info.nvar, h.info.ncon, h.info.nnzj, h.info.nnzh) == (orc.nvar, orc.ncon, orc.nnzj, orc.nnzh)

    def close(a, b, what):
        s = max(float(np.abs(b).max()), 1e-300)
        assert float(np.abs(a - b).max()) <= 1e-12 * s, what

    close(np.array([h.obj(x)]), np.array([orc.obj(x)]), "obj")
    close(h.grad(x), orc.grad(x), "grad")
    close(h.cons(x), orc.cons(x), "cons")
    close(h.jprod(x, v), orc.jprod(x, v), "jprod")
    close(h.jtprod(x, lam), orc.jtprod(x, lam), "jtprod")
    close(h.hprod(x, lam, v, 0.37), orc.hprod(x, v, lam, 0.37), "hprod")
    close(h.jac_coord(x), orc.jac_coord(x), "jac_coord")
    close(h.hess_coord(x, lam, 0.37), orc.hess_coord(x, lam, 0.37), "hess_coord")
    for hess in (False, True):
        r, c = h.structure(hess)
        ro, co = orc.hess_structure() if hess else orc.jac_structure()
        assert np.array_equal(r, ro) and np.array_equal(c, co)
    h.close()


@pytest.mark.gpu
def test_sharded_handle_at_a_template_active_size_equals_the_single_device_model(pkg, cuda_lib):
    """pdenlp_sharded_* (one process, one slab per device, host vectors) on slabs large enough for the cell-granular
    template path (>= 4096 tiles per slab): global jac_coord! / hess_coord! streams equal the single-device model's —
    FE segments bit for bit, the dense kappa blocks (host-side sums of slab partials) to round-off."""
    import torch
    from pdenlp_b200.capi import ShardedHandle
    from pdenlp_b200.flatten import flatten
    from pdenlp_b200.model import GridapPDENLPModel
    ndev = min(2, torch.cuda.device_count())
    spec = pkg.poissonmixed(64 if ndev == 1 else 80, 3, True)   # 262 144 / 512 000 cells
    fm = flatten(spec)
    h = ShardedHandle(fm, ndev)
    d = GridapPDENLPModel(spec, device="cuda:0", flat=fm)
    try:
        I = d.handle.info
        g = np.random.default_rng(13)
        x = g.uniform(-1, 1, d.meta.nvar); x[: spec.nparam] = 1 + 0.1 * g.uniform(-1, 1, spec.nparam)
        lam = g.uniform(-1, 1, d.meta.ncon)
        xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lam).cuda()
        nk = int(I.nnzj_k)
        jh, jd = h.jac_coord(x), d.jac_coord(xd).cpu().numpy()
        assert np.array_equal(jh[nk:], jd[nk:]) and np.allclose(jh[:nk], jd[:nk], rtol=1e-11, atol=1e-13 * np.abs(jd[:nk]).max())
        hh, hd = h.hess_coord(x, lam, 0.37), d.hess_coord(xd, ld, obj_weight=0.37).cpu().numpy()
        ko, kc, fo = int(I.nnzh_k_obj), int(I.nnzh_k_con), int(I.nnzh_fe_obj)
        assert np.array_equal(hh[ko: ko + fo], hd[ko: ko + fo]) and np.array_equal(hh[ko + fo + kc:], hd[ko + fo + kc:])
        assert np.allclose(hh, hd, rtol=1e-11, atol=1e-13 * np.abs(hd).max())
    finally:
        h.close(); d.close()
