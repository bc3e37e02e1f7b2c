"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): structure bit-exact; values within a relative
tolerance of 1e-12 in FP64 — element-wise for every entry above 1e-3 of its segment's
scale, segment-relative for the rest (tests/parity.py).
"""
import numpy as np
import pytest
import torch

from parity import RTOL, assert_parity

pytestmark = pytest.mark.gpu


def _close(a, b, what):
    assert_parity(a, b, what)


def _cases(pkg):
    return {
        "membrane3": lambda: pkg.controlelasticmembrane(3),
        "membrane8": lambda: pkg.controlelasticmembrane(8),
        "membrane37": lambda: pkg.controlelasticmembrane(37),
        "membrane100": lambda: pkg.controlelasticmembrane(100),
        "pb_l2_9": lambda: pkg.poissonboltzman2d(9),
        "pb_h1_33": lambda: pkg.poissonboltzman2d(33, "H1"),
        "burger_lin_70": lambda: pkg.burger1d(70),
        "burger_nl_70": lambda: pkg.burger1d_param(70),
        "kappa2d_5": lambda: pkg.poissonmixed(5),
        "kappa2d_u_6": lambda: pkg.poissonmixed(6, 2, True),
        "kappa3d_u_4": lambda: pkg.poissonmixed(4, 3, True),
        # multi-field state / control spaces (block-ordered COO)
        "controlsir_70": lambda: pkg.controlsir(70),
        "cellincrease_45": lambda: pkg.cellincrease(45),
        "dynamicsir_100": lambda: pkg.dynamicsir(100),
        "dynamicsir_5": lambda: pkg.dynamicsir(5),
        # quadrature degree 2-3: four / two Gauss points per cell (reference-tensor path for the affine membrane,
        # per-point accumulation for the nonlinear ones)
        "membrane20_deg2": lambda: pkg.controlelasticmembrane(20, degree=2),
        "membrane_q1_33_deg3": lambda: pkg.controlelasticmembrane(33, degree=3, state_order=1),
        "pb_l2_21_deg2": lambda: pkg.poissonboltzman2d(21, degree=2),
        "burger_lin_90_deg2": lambda: pkg.burger1d(90, degree=2),
        "burger_nl_90_deg3": lambda: pkg.burger1d_param(90, degree=3),
        "poissonmixed2_17": lambda: pkg.poissonmixed2(17),   # inde = false: kappa-y cross rows in the objective block
        # 3-D Q1/Q1
        "membrane3d_6": lambda: pkg.controlelasticmembrane(6, state_order=1, dim=3),
        "membrane3d_5_deg2": lambda: pkg.controlelasticmembrane(5, state_order=1, dim=3, degree=2),
        "pb3d_l2_6": lambda: pkg.poissonboltzman2d(6, dim=3),
        "pb_q2_19": lambda: pkg.poissonboltzman2d(19, "H1", state_order=2),
        "pb_q2_11_deg2": lambda: pkg.poissonboltzman2d(11, state_order=2, degree=2),   # per-point accumulation in shared memory
        # NoFETerm objectives: no objective FE block in the Hessian
        "poissonparam_24": lambda: pkg.poissonparam(24),
        "sis_80": lambda: pkg.sis(80),
    }


CASE_NAMES = ["membrane3", "membrane8", "membrane37", "membrane100", "pb_l2_9", "pb_h1_33", "burger_lin_70",
              "burger_nl_70", "kappa2d_5", "kappa2d_u_6", "kappa3d_u_4",
              "controlsir_70", "cellincrease_45", "dynamicsir_100", "dynamicsir_5", "poissonparam_24", "sis_80",
              "membrane20_deg2", "membrane_q1_33_deg3", "pb_l2_21_deg2", "burger_lin_90_deg2", "burger_nl_90_deg3",
              "poissonmixed2_17", "membrane3d_6", "membrane3d_5_deg2", "pb3d_l2_6",
              "pb_q2_19", "pb_q2_11_deg2"]


@pytest.fixture(scope="module", params=CASE_NAMES)
def case(request, pkg, oracle_mod, cuda_lib):
    from pdenlp_b200.model import GridapPDENLPModel
    spec = _cases(pkg)[request.param]()
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    orc = oracle_mod.Oracle(spec)
    rng = np.random.default_rng(1998)
    x = rng.uniform(-1, 1, nlp.meta.nvar)
    if spec.nparam:
        x[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
    lam = np.random.default_rng(1999).uniform(-1, 1, nlp.meta.ncon)
    v = np.random.default_rng(2000).uniform(-1, 1, nlp.meta.nvar)
    yield dict(name=request.param, spec=spec, nlp=nlp, orc=orc, x=x, lam=lam, v=v)
    nlp.close()


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_sizes(case):
    nlp, orc = case["nlp"], case["orc"]
    assert (nlp.meta.nvar, nlp.meta.ncon, nlp.meta.nnzj, nlp.meta.nnzh) == (orc.nvar, orc.ncon, orc.nnzj, orc.nnzh)
    assert nlp.pdemeta.nnzh_obj == orc.nnzh_obj


def test_structure_bit_exact(case):
    nlp, orc = case["nlp"], case["orc"]
    r, c = nlp.jac_structure()
    ro, co = orc.jac_structure()
    assert np.array_equal(r.cpu().numpy(), ro) and np.array_equal(c.cpu().numpy(), co)
    r, c = nlp.hess_structure()
    ro, co = orc.hess_structure()
    assert np.array_equal(r.cpu().numpy(), ro) and np.array_equal(c.cpu().numpy(), co)


@pytest.mark.parametrize("xkind", ["rand", "zeros", "ones"])
def test_jac_coord(case, xkind):
    nlp, orc = case["nlp"], case["orc"]
    x = {"rand": case["x"], "zeros": np.zeros_like(case["x"]), "ones": np.ones_like(case["x"])}[xkind]
    _close(nlp.jac_coord(_d(x)).cpu().numpy(), orc.jac_coord(x), "jac_coord")


@pytest.mark.parametrize("ow", [1.0, 0.0, 0.37])
def test_hess_coord_lagrangian(case, ow):
    nlp, orc, x, lam = case["nlp"], case["orc"], case["x"], case["lam"]
    got = nlp.hess_coord(_d(x), _d(lam), obj_weight=ow).cpu().numpy()
    ref = orc.hess_coord(x, lam, ow)
    no = orc.nnzh_obj
    _close(got[:no], ref[:no], "hess_coord obj part")
    _close(got[no:], ref[no:], "hess_coord cons part")


def test_hess_coord_objective_only(case):
    nlp, orc, x = case["nlp"], case["orc"], case["x"]
    got = nlp.hess_coord(_d(x), obj_weight=0.37).cpu().numpy()
    ref = orc.hess_coord(x, None, 0.37)
    _close(got, ref, "hess_coord(x)")
    assert np.all(got[orc.nnzh_obj:] == 0.0)  # tail zero-filled, GridapPDENLPModel.jl:295-297


def test_obj_grad_cons(case):
    nlp, orc, x = case["nlp"], case["orc"], case["x"]
    f, fo = nlp.obj(_d(x)), orc.obj(x)
    assert abs(f - fo) <= RTOL * max(abs(fo), 1e-300)
    _close(nlp.grad(_d(x)).cpu().numpy(), orc.grad(x), "grad")
    _close(nlp.cons(_d(x)).cpu().numpy(), orc.cons(x), "cons")


def test_products(case):
    nlp, orc, x, lam, v = case["nlp"], case["orc"], case["x"], case["lam"], case["v"]
    _close(nlp.jprod(_d(x), _d(v)).cpu().numpy(), orc.jprod(x, v), "jprod")
    _close(nlp.jtprod(_d(x), _d(lam)).cpu().numpy(), orc.jtprod(x, lam), "jtprod")
    _close(nlp.hprod(_d(x), _d(lam), _d(v), obj_weight=0.37).cpu().numpy(), orc.hprod(x, v, lam, 0.37), "hprod(x,y,v)")
    _close(nlp.hprod(_d(x), _d(v), obj_weight=0.37).cpu().numpy(), orc.hprod(x, v, None, 0.37), "hprod(x,v)")
    assert torch.all(nlp.hprod(_d(x), _d(v), obj_weight=0.0) == 0)  # GridapPDENLPModel.jl:313-316


def test_host_memspace_matches_device(case, pkg):
    """PDENLP_MEM_HOST: numpy in/out through the library's staging copies."""
    from pdenlp_b200.model import GridapPDENLPModel
    if case["name"] not in ("membrane8", "burger_nl_70", "dynamicsir_100"):
        pytest.skip("one small case per family is enough")
    nlp_h = GridapPDENLPModel(case["spec"], device="cuda:0", memspace="host")
    x, lam, v = case["x"], case["lam"], case["v"]
    # jtprod! FIRST: its direction has length ncon < nvar, and the staging buffer of the directions is shared with
    # jprod!/hprod! (length nvar) — it must grow (round-1 bug: sized by first use, later overflowed)
    nlp_d = case["nlp"]
    assert nlp_h.meta.nvar > nlp_h.meta.ncon
    big = 1e3 * np.ones_like(v)  # a direction whose overflowing tail would be visible in neighbouring arrays
    tol = dict(rtol=1e-12, atol=1e-13)
    assert np.allclose(nlp_h.jtprod(x, lam), nlp_d.jtprod(_d(x), _d(lam)).cpu().numpy(), **tol)
    assert np.allclose(nlp_h.jprod(x, big), nlp_d.jprod(_d(x), _d(big)).cpu().numpy(), rtol=1e-12, atol=1e-9)
    assert np.allclose(nlp_h.hprod(x, lam, v, obj_weight=0.37), nlp_d.hprod(_d(x), _d(lam), _d(v), obj_weight=0.37).cpu().numpy(), **tol)
    assert np.allclose(nlp_h.jtprod(x, lam), nlp_d.jtprod(_d(x), _d(lam)).cpu().numpy(), **tol)
    assert np.allclose(nlp_h.grad(x), nlp_d.grad(_d(x)).cpu().numpy(), **tol)
    jd = case["nlp"].jac_coord(_d(x)).cpu().numpy()
    hd = case["nlp"].hess_coord(_d(x), _d(lam), obj_weight=0.5).cpu().numpy()
    assert np.array_equal(nlp_h.jac_coord(x), jd)
    assert np.array_equal(nlp_h.hess_coord(x, lam, obj_weight=0.5), hd)
    r, c = nlp_h.jac_structure()
    ro, co = case["orc"].jac_structure()
    assert np.array_equal(r, ro) and np.array_equal(c, co)
    nlp_h.close()


def test_host_memspace_slab_stages_only_its_dof_ranges(pkg):
    """A cell slab in PDENLP_MEM_HOST stages only the dof ranges its cells touch (pdenlp_input_ranges); every callback
    must equal the device-memspace slab model — with NaNs in the entries of x / lambda / v the slab never reads, so a
    read outside the staged ranges (or a wrong range) shows up as a NaN."""
    from pdenlp_b200.model import GridapPDENLPModel
    for spec in (pkg.controlelasticmembrane(12), pkg.poissonmixed(12, 2, True)):
        nc = spec.model.ncells
        rng = (32 * ((nc // 3) // 32), 32 * ((2 * nc // 3) // 32))
        nd = GridapPDENLPModel(spec, device="cuda:0", cell_range=rng)
        nh = GridapPDENLPModel(spec, device="cuda:0", cell_range=rng, memspace="host")
        ylo, yhi, ulo, uhi = nh.handle.input_ranges()
        n, m, p = nh.meta.nvar, nh.meta.ncon, spec.nparam
        assert 0 <= ylo < yhi <= m and (yhi - ylo) < m, "the slab must touch a strict sub-range of the state dofs"
        assert nh.staged_input_bytes(("x",)) == 8 * (p + (yhi - ylo) + (uhi - ulo)) < 8 * n
        g = np.random.default_rng(5)
        x, lam, v = g.uniform(-1, 1, n), g.uniform(-1, 1, m), g.uniform(-1, 1, n)
        if p:
            x[:p] = 1 + 0.1 * g.uniform(-1, 1, p)
        xh, lh, vh = x.copy(), lam.copy(), v.copy()
        keep = np.zeros(n, bool); keep[:p] = True; keep[p + ylo: p + yhi] = True; keep[p + m + ulo: p + m + uhi] = True
        xh[~keep] = np.nan; vh[~keep] = np.nan
        kc = np.zeros(m, bool); kc[ylo:yhi] = True
        lh[~kc] = np.nan
        tol = dict(rtol=1e-12, atol=1e-13)
        # (FE segments are bitwise reproducible; the dense kappa blocks are sums over cells through atomics: round-off)
        same = np.array_equal if p == 0 else (lambda a, b: np.allclose(a, b, **tol))
        assert same(nh.jac_coord(xh), nd.jac_coord(_d(x)).cpu().numpy())
        assert same(nh.hess_coord(xh, lh, obj_weight=0.5), nd.hess_coord(_d(x), _d(lam), obj_weight=0.5).cpu().numpy())
        assert np.allclose(nh.grad(xh), nd.grad(_d(x)).cpu().numpy(), **tol)
        assert np.allclose(nh.cons(xh), nd.cons(_d(x)).cpu().numpy(), **tol)
        assert np.allclose(nh.jprod(xh, vh), nd.jprod(_d(x), _d(v)).cpu().numpy(), **tol)
        assert np.allclose(nh.jtprod(xh, lh), nd.jtprod(_d(x), _d(lam)).cpu().numpy(), **tol)
        assert np.allclose(nh.hprod(xh, lh, vh, obj_weight=0.37), nd.hprod(_d(x), _d(lam), _d(v), obj_weight=0.37).cpu().numpy(), **tol)
        assert abs(nh.obj(xh) - nd.obj(_d(x))) <= 1e-12 * max(1.0, abs(nd.obj(_d(x))))
        nh.close(); nd.close()


def test_deterministic_mode_is_bitwise_reproducible_and_equal_to_the_atomic_path(pkg):
    """PDENLP_FLAG_DETERMINISTIC: coloured launches with plain read-modify-write accumulation.  Every dof-indexed output
    and the dense kappa blocks must (a) agree with the default (atomic) kernels to round-off and (b) be bitwise identical
    over repeated evaluations — also against a second model of the same problem (fresh colouring, fresh buffers)."""
    from pdenlp_b200.model import GridapPDENLPModel
    for spec in (pkg.controlelasticmembrane(40), pkg.poissonmixed(24, 2, True), pkg.burger1d_param(3000), pkg.poissonboltzman2d(48)):
        nd = GridapPDENLPModel(spec, device="cuda:0")
        n1 = GridapPDENLPModel(spec, device="cuda:0", deterministic=True)
        n2 = GridapPDENLPModel(spec, device="cuda:0", deterministic=True)
        n, m, p = nd.meta.nvar, nd.meta.ncon, spec.nparam
        g = np.random.default_rng(7)
        x, lam, v = g.uniform(-1, 1, n), g.uniform(-1, 1, m), g.uniform(-1, 1, n)
        if p:
            x[:p] = 1 + 0.1 * g.uniform(-1, 1, p)
        x, lam, v = _d(x), _d(lam), _d(v)
        calls = {
            "grad": lambda q: q.grad(x), "cons": lambda q: q.cons(x), "jprod": lambda q: q.jprod(x, v),
            "jtprod": lambda q: q.jtprod(x, lam), "hprod": lambda q: q.hprod(x, lam, v, obj_weight=0.37),
            "jac_coord": lambda q: q.jac_coord(x), "hess_coord": lambda q: q.hess_coord(x, lam, obj_weight=0.37),
            "objgrad": lambda q: q.objgrad(x)[1],
        }
        for name, fn in calls.items():
            ref = fn(nd).cpu().numpy()
            a = fn(n1).cpu().numpy()
            _close(a, ref, f"{spec.name} {name} deterministic vs atomic")
            for _ in range(3):
                assert np.array_equal(fn(n1).cpu().numpy(), a), f"{spec.name} {name}: not reproducible"
            assert np.array_equal(fn(n2).cpu().numpy(), a), f"{spec.name} {name}: differs between two models"
        assert abs(n1.obj(x) - nd.obj(x)) <= 1e-12 * max(1.0, abs(nd.obj(x)))
        for q in (nd, n1, n2):
            q.close()


def test_unaligned_guarded_outputs(case):
    """Outputs that are only 8 B aligned (a view starting at an odd element) with sentinel guards on
    both sides: exercises the head/tail handling of the 16 B-aligned bulk stores and catches any
    out-of-bounds write (compute-sanitizer is closed on this pool)."""
    nlp, x, lam = case["nlp"], case["x"], case["lam"]
    xd, ld = _d(x), _d(lam)
    jref = nlp.jac_coord(xd)
    href = nlp.hess_coord(xd, ld, obj_weight=0.37)
    for off in (1, 2, 3):
        for ref, call in ((jref, lambda out: nlp.jac_coord_(xd, out)),
                          (href, lambda out: nlp.hess_coord_(xd, ld, out, obj_weight=0.37))):
            n = ref.numel()
            big = torch.full((n + 16,), float("nan"), dtype=torch.float64, device="cuda")
            view = big[off: off + n]
            assert view.data_ptr() % 16 == (8 * off) % 16
            call(view)
            torch.cuda.synchronize()
            if case["spec"].nparam == 0:
                assert torch.equal(view, ref)
            else:  # the dense kappa blocks are accumulated with atomics: equal up to summation order
                assert torch.allclose(view, ref, rtol=1e-12, atol=1e-13 * float(ref.abs().max()))
            assert torch.isnan(big[:off]).all() and torch.isnan(big[off + n:]).all()


@pytest.mark.parametrize("which", ["penalizedpoisson", "basicunconstrained", "torebrachistochrone"])
def test_unconstrained_models(which, pkg, oracle_mod, cuda_lib):
    """The reference's unconstrained constructor (src/additional_constructors.jl:1-66): ncon = 0,
    empty Jacobian, Hessian = objective segment; values against the oracle."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = getattr(pkg, which)(24)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    orc = oracle_mod.Oracle(spec)
    no = orc.nnzh_obj
    assert (nlp.meta.nvar, nlp.meta.ncon, nlp.meta.nnzj, nlp.meta.nnzh) == (orc.nvar, 0, 0, no)
    rng = np.random.default_rng(1998)
    x = rng.uniform(-1, 1, orc.nvar); v = rng.uniform(-1, 1, orc.nvar)
    f, fo = nlp.obj(_d(x)), orc.obj(x)
    assert abs(f - fo) <= RTOL * abs(fo)
    _close(nlp.grad(_d(x)).cpu().numpy(), orc.grad(x), "grad")
    _close(nlp.hess_coord(_d(x), obj_weight=0.37).cpu().numpy(), orc.hess_coord(x, None, 0.37)[:no], "hess_coord")
    _close(nlp.hprod(_d(x), _d(v), obj_weight=0.37).cpu().numpy(), orc.hprod(x, v, None, 0.37), "hprod")
    r, c = nlp.hess_structure(); ro, co = orc.hess_structure()
    assert np.array_equal(r.cpu().numpy(), ro[:no]) and np.array_equal(c.cpu().numpy(), co[:no])
    empty = torch.empty(0, dtype=torch.float64, device="cuda")
    assert nlp.cons(_d(x)).numel() == 0 and nlp.jac_coord(_d(x)).numel() == 0 and nlp.jprod(_d(x), _d(v)).numel() == 0
    assert torch.all(nlp.jtprod(_d(x), empty) == 0)
    nlp.close()


def test_callbacks_follow_the_current_torch_stream(pkg, cuda_lib):
    """pdenlp_set_stream: every callback runs on torch's current stream, so producer / consumer kernels
    enqueued on a side stream are ordered with it without host synchronisation."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = pkg.controlelasticmembrane(64)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    rng = np.random.default_rng(3)
    x0 = _d(rng.uniform(-1, 1, nlp.meta.nvar)); lam = _d(rng.uniform(-1, 1, nlp.meta.ncon))
    ref_j = nlp.jac_coord(2.0 * x0); ref_h = nlp.hess_coord(2.0 * x0, lam, obj_weight=0.5); ref_g = nlp.grad(2.0 * x0)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        x = x0 * 2.0                         # producer on the side stream
        j = nlp.jac_coord(x); h = nlp.hess_coord(x, lam, obj_weight=0.5); g = nlp.grad(x)
        js = j.clone(); hs = h.clone()       # consumers on the side stream
    side.synchronize()
    assert torch.equal(js, ref_j) and torch.equal(hs, ref_h)
    assert torch.allclose(g, ref_g, rtol=1e-13, atol=1e-15)
    nlp.close()


@pytest.mark.parametrize("case", ["membrane48", "membrane512_templates", "kappa3d64_templates_aux_stream", "burger_param_aux_stream"])
def test_callbacks_can_be_captured_in_a_cuda_graph(pkg, cuda_lib, case):
    """Device-memspace coordinate / vector callbacks enqueue kernels only (no host synchronisation, no host
    copies), so a solver can capture an iteration in a CUDA graph and replay it with new x / lambda contents.
    The larger cases run the cell-granular template path and fork / join the model's auxiliary stream inside the
    capture (event record + wait: the auxiliary stream joins the capture and is joined back before the callback returns)."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = {"membrane48": lambda: pkg.controlelasticmembrane(48),
            "membrane512_templates": lambda: pkg.controlelasticmembrane(512),
            "kappa3d64_templates_aux_stream": lambda: pkg.poissonmixed(64, 3, True),
            "burger_param_aux_stream": lambda: pkg.burger1d_param(100000)}[case]()
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    rng = np.random.default_rng(4)
    def _x():
        v = rng.uniform(-1, 1, nlp.meta.nvar)
        if spec.nparam:
            v[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
        return v
    x = _d(_x()); lam = _d(rng.uniform(-1, 1, nlp.meta.ncon))
    jv = torch.empty(nlp.meta.nnzj, dtype=torch.float64, device="cuda")
    hv = torch.empty(nlp.meta.nnzh, dtype=torch.float64, device="cuda")
    g = torch.empty(nlp.meta.nvar, dtype=torch.float64, device="cuda")
    c = torch.empty(nlp.meta.ncon, dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):   # warm-up outside the capture (function attributes, occupancy queries)
        nlp.jac_coord_(x, jv); nlp.hess_coord_(x, lam, hv, obj_weight=0.5); nlp.grad_(x, g); nlp.cons_(x, c)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        nlp.jac_coord_(x, jv); nlp.hess_coord_(x, lam, hv, obj_weight=0.5); nlp.grad_(x, g); nlp.cons_(x, c)
    x2 = _d(_x()); lam2 = _d(rng.uniform(-1, 1, nlp.meta.ncon))
    x.copy_(x2); lam.copy_(lam2)
    jv.fill_(float("nan")); hv.fill_(float("nan"))
    graph.replay()
    torch.cuda.synchronize()
    jr, hr = nlp.jac_coord(x2), nlp.hess_coord(x2, lam2, obj_weight=0.5)
    if spec.nparam == 0:
        assert torch.equal(jv, jr) and torch.equal(hv, hr)
    else:  # (the dense kappa blocks are atomic sums: equal up to round-off; the FE segments bit for bit)
        I = nlp.handle.info
        assert torch.equal(jv[int(I.nnzj_k):], jr[int(I.nnzj_k):])
        assert torch.allclose(jv, jr, rtol=1e-12, atol=1e-14 * float(jr.abs().max()))
        assert torch.allclose(hv, hr, rtol=1e-12, atol=1e-14 * float(hr.abs().max()))
        assert not torch.isnan(hv).any() and not torch.isnan(jv).any()
    assert torch.allclose(g, nlp.grad(x2), rtol=1e-13, atol=1e-15) and torch.allclose(c, nlp.cons(x2), rtol=1e-13, atol=1e-15)
    nlp.close()


@pytest.mark.parametrize("case", ["burger1d_param", "poissonmixed2d", "poissonmixed2"])
def test_graph_replay_of_small_kappa_models_matches_plain_calls(pkg, cuda_lib, case):
    """Small models with kappa blocks replay jac_coord! / hess_coord! as one CUDA graph from the second call with the
    same buffers on (include/pdenlp_cuda.h: pdenlp_graph_replays).  The replays must read the CURRENT contents of x and
    lambda and equal a model that never replays (fresh buffers every call): FE segments bit for bit, kappa blocks
    (atomic sums) to round-off; a different obj_weight or buffer is a different graph."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = {"burger1d_param": lambda: pkg.burger1d_param(20000), "poissonmixed2d": lambda: pkg.poissonmixed(40, 2, True),
            "poissonmixed2": lambda: pkg.poissonmixed2(40)}[case]()
    a = GridapPDENLPModel(spec, device="cuda:0")
    b = GridapPDENLPModel(spec, device="cuda:0")
    I = a.handle.info
    rng = np.random.default_rng(11)
    x = torch.empty(a.meta.nvar, dtype=torch.float64, device="cuda")
    lam = torch.empty(a.meta.ncon, dtype=torch.float64, device="cuda")
    jv = torch.empty(a.meta.nnzj, dtype=torch.float64, device="cuda")
    hv = torch.empty(a.meta.nnzh, dtype=torch.float64, device="cuda")
    nk = int(I.nnzj_k)
    for it in range(5):
        xh = rng.uniform(-1, 1, a.meta.nvar); xh[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
        lh = rng.uniform(-1, 1, a.meta.ncon)
        x.copy_(torch.from_numpy(xh)); lam.copy_(torch.from_numpy(lh))
        ow = 0.37 if it % 2 else 1.0
        jv.fill_(float("nan")); hv.fill_(float("nan"))
        a.jac_coord_(x, jv); a.hess_coord_(x, lam, hv, obj_weight=ow)
        jr = b.jac_coord(x.clone()); hr = b.hess_coord(x.clone(), lam.clone(), obj_weight=ow)   # fresh buffers: never replayed
        torch.cuda.synchronize()
        assert torch.equal(jv[nk:], jr[nk:])
        assert torch.allclose(jv, jr, rtol=1e-12, atol=1e-14 * float(jr.abs().max()))
        assert torch.allclose(hv, hr, rtol=1e-12, atol=1e-14 * float(hr.abs().max()))
        assert not torch.isnan(jv).any() and not torch.isnan(hv).any()
    # (b may replay too when the caching allocator hands a freed buffer back at the same address: equally valid)
    assert a.handle.graph_replays() >= 5, a.handle.graph_replays()
    a.close(); b.close()
