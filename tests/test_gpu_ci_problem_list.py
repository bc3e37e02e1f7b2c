"""The reference's CI loop (test/runtests.jl:15-89) replayed on the CUDA path.

For every problem of `pde_problems` (CI list + the local-only ones that are built here), at n = 3 as in
runtests.jl:15, the three NLPModelsTest families the reference runs:
  * consistent_nlps      -- obj/grad, hess vs hess_coord vs hprod vs hess_op, cons, jac vs jac_coord vs
                            jprod/jtprod vs jac_op agree with each other (rtol 1e-8) at deterministic points
  * check_nlp_dimensions -- every API function raises DimensionError on wrong-length vectors
  * view_subarray_nlp    -- calls with views of larger arrays give the plain-array result
and, beyond what the reference checks, values against the CPU oracle (1e-12).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu

N = 3
PROBLEMS = {
    # CI list (test/runtests.jl:40-49)
    "BURGER1D": lambda p: p.burger1d(N),
    "BURGER1D_PARAM": lambda p: p.burger1d_param(N),
    "CONTROLSIR": lambda p: p.controlsir(N),
    "PENALIZEDPOISSON": lambda p: p.penalizedpoisson(N),
    "POISSONMIXED": lambda p: p.poissonmixed(N),
    "POISSONPARAM": lambda p: p.poissonparam(N),
    "TOREBRACHISTOCHRONE": lambda p: p.torebrachistochrone(N),
    "CONTROLELASTICMEMBRANE": lambda p: p.controlelasticmembrane(N, affine=True),
    # local-only list (test/runtests.jl:22-37)
    "CELLINCREASE": lambda p: p.cellincrease(N),
    "SIS": lambda p: p.sis(N),
    "DYNAMICSIR": lambda p: p.dynamicsir(N),
    "BASICUNCONSTRAINED": lambda p: p.basicunconstrained(N),
    "POISSONMIXED2": lambda p: p.poissonmixed2(N),
}


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _h(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module", params=sorted(PROBLEMS))
def prob(request, pkg, cuda_lib):
    from pdenlp_b200.model import GridapPDENLPModel
    spec = PROBLEMS[request.param](pkg)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    yield request.param, spec, nlp
    nlp.close()


def _points(n):
    i = np.arange(1, n + 1)
    # NLPModelsTest evaluates at 10*[-(-1)^i]-style points; scaled down to keep sinh/cos integrands tame
    return [0.1 * (-(-1.0) ** i), np.zeros(n), np.random.default_rng(7).uniform(-1, 1, n)]


def _dense_sym(nlp, vals, nnz=None):
    r, c = nlp.hess_structure()
    r, c, vals = _h(r), _h(c), _h(vals)
    H = sp.coo_matrix((vals, (r - 1, c - 1)), shape=(nlp.meta.nvar,) * 2).toarray()
    return H + np.tril(H, -1).T


def test_consistency(prob):
    name, spec, nlp = prob
    n, m = nlp.meta.nvar, nlp.meta.ncon
    rng = np.random.default_rng(11)
    for x in _points(n):
        xd = _d(x)
        y = rng.uniform(-1, 1, m); v = rng.uniform(-1, 1, n); w = rng.uniform(-1, 1, m)
        g = _h(nlp.grad(xd))
        # objective-only Hessian: hess_coord == hprod == hess_op
        H = _dense_sym(nlp, nlp.hess_coord(xd, obj_weight=0.7))
        assert np.allclose(_h(nlp.hprod(xd, _d(v), obj_weight=0.7)), H @ v, rtol=1e-8, atol=1e-10)
        assert np.allclose(_h(nlp.hess_op(xd, obj_weight=0.7)(_d(v))), H @ v, rtol=1e-8, atol=1e-10)
        # gradient against a directional finite difference of obj (gradient_check analogue)
        eps = 1e-6
        fd = (nlp.obj(_d(x + eps * v)) - nlp.obj(_d(x - eps * v))) / (2 * eps)
        assert abs(fd - g @ v) <= 1e-6 * max(1.0, abs(g @ v))
        if m == 0:
            assert nlp.meta.nnzj == 0 and nlp.cons(xd).numel() == 0
            continue
        # Lagrangian Hessian
        HL = _dense_sym(nlp, nlp.hess_coord(xd, _d(y), obj_weight=0.7))
        assert np.allclose(_h(nlp.hprod(xd, _d(y), _d(v), obj_weight=0.7)), HL @ v, rtol=1e-8, atol=1e-10)
        assert np.allclose(_h(nlp.hess_op(xd, _d(y), obj_weight=0.7)(_d(v))), HL @ v, rtol=1e-8, atol=1e-10)
        # Jacobian: jac_coord == jprod == jtprod == jac_op, and cons against a finite difference
        r, c = nlp.jac_structure()
        J = sp.coo_matrix((_h(nlp.jac_coord(xd)), (_h(r) - 1, _h(c) - 1)), shape=(m, n)).toarray()
        assert np.allclose(_h(nlp.jprod(xd, _d(v))), J @ v, rtol=1e-8, atol=1e-10)
        assert np.allclose(_h(nlp.jtprod(xd, _d(w))), J.T @ w, rtol=1e-8, atol=1e-10)
        jop, jtop = nlp.jac_op(xd)
        assert np.allclose(_h(jop(_d(v))), J @ v, rtol=1e-8, atol=1e-10)
        assert np.allclose(_h(jtop(_d(w))), J.T @ w, rtol=1e-8, atol=1e-10)
        cfd = (_h(nlp.cons(_d(x + eps * v))) - _h(nlp.cons(_d(x - eps * v)))) / (2 * eps)
        assert np.allclose(cfd, J @ v, rtol=1e-5, atol=1e-6)


def test_values_against_oracle(prob, oracle_mod):
    name, spec, nlp = prob
    if getattr(spec, "affine", False):
        pytest.skip("linear-API values are covered by tests/test_gpu_linear_api.py")
    orc = oracle_mod.Oracle(spec)
    n, m = nlp.meta.nvar, nlp.meta.ncon
    x = _points(n)[2]; xd = _d(x)
    y = np.random.default_rng(12).uniform(-1, 1, m)

    from parity import is_close as close  # element-wise + segment-relative 1e-12 (tests/parity.py)

    assert abs(nlp.obj(xd) - orc.obj(x)) <= 1e-12 * max(abs(orc.obj(x)), 1e-300)
    assert close(_h(nlp.grad(xd)), orc.grad(x))
    no = orc.nnzh_obj
    assert close(_h(nlp.hess_coord(xd, obj_weight=0.7))[:no], orc.hess_coord(x, None, 0.7)[:no])
    if m:
        assert close(_h(nlp.cons(xd)), orc.cons(x)) and close(_h(nlp.jac_coord(xd)), orc.jac_coord(x))
        assert close(_h(nlp.hess_coord(xd, _d(y), obj_weight=0.7)), orc.hess_coord(x, y, 0.7))
        r, c = nlp.jac_structure(); ro, co = orc.jac_structure()
        assert np.array_equal(_h(r), ro) and np.array_equal(_h(c), co)
    r, c = nlp.hess_structure(); ro, co = orc.hess_structure()
    k = nlp.meta.nnzh
    assert np.array_equal(_h(r), ro[:k]) and np.array_equal(_h(c), co[:k])


def test_check_nlp_dimensions(prob):
    from pdenlp_b200.model import DimensionError
    name, spec, nlp = prob
    n, m, nnzj, nnzh = nlp.meta.nvar, nlp.meta.ncon, nlp.meta.nnzj, nlp.meta.nnzh
    z = lambda k: torch.zeros(max(k, 0), dtype=torch.float64, device="cuda")
    zi = lambda k: torch.zeros(max(k, 0), dtype=torch.int64, device="cuda")
    x = z(n)
    bad = [
        lambda: nlp.obj(z(n + 1)), lambda: nlp.grad_(z(n + 1), z(n)), lambda: nlp.grad_(x, z(n + 1)),
        lambda: nlp.hess_coord_(z(n + 1), z(nnzh)), lambda: nlp.hess_coord_(x, z(nnzh + 1)),
        lambda: nlp.hprod_(x, z(n + 1), z(n)), lambda: nlp.hprod_(x, z(n), z(n + 1)),
        lambda: nlp.hess_structure_(zi(nnzh + 1), zi(nnzh)), lambda: nlp.hess_structure_(zi(nnzh), zi(nnzh + 1)),
        lambda: nlp.cons_(z(n + 1), z(m)), lambda: nlp.cons_(x, z(m + 1)),
        lambda: nlp.jac_coord_(z(n + 1), z(nnzj)), lambda: nlp.jac_coord_(x, z(nnzj + 1)),
        lambda: nlp.jprod_(x, z(n + 1), z(m)), lambda: nlp.jprod_(x, z(n), z(m + 1)),
        lambda: nlp.jtprod_(x, z(m + 1), z(n)), lambda: nlp.jtprod_(x, z(m), z(n + 1)),
        lambda: nlp.jac_structure_(zi(nnzj + 1), zi(nnzj)),
        lambda: nlp.hess_coord_(x, z(m + 1), z(nnzh)), lambda: nlp.hprod_(x, z(m + 1), z(n), z(n)),
    ]
    for f in bad:
        with pytest.raises(DimensionError):
            f()


def test_view_subarray(prob):
    name, spec, nlp = prob
    n, m = nlp.meta.nvar, nlp.meta.ncon
    rng = np.random.default_rng(5)
    big_x = _d(rng.uniform(-1, 1, n + 7)); big_v = _d(rng.uniform(-1, 1, n + 5)); big_y = _d(rng.uniform(-1, 1, m + 3))
    xv, vv, yv = big_x[3: 3 + n], big_v[2: 2 + n], big_y[1: 1 + m]      # views at 8 B-aligned offsets
    x, v, y = xv.clone(), vv.clone(), yv.clone()
    tol = dict(rtol=1e-12, atol=1e-13)  # dof-indexed outputs are scatter-added with atomics
    assert abs(nlp.obj(xv) - nlp.obj(x)) <= 1e-13 * max(1.0, abs(nlp.obj(x)))
    assert torch.allclose(nlp.grad(xv), nlp.grad(x), **tol)
    assert torch.allclose(nlp.hprod(xv, vv, obj_weight=0.5), nlp.hprod(x, v, obj_weight=0.5), **tol)
    assert torch.allclose(nlp.hess_coord(xv, obj_weight=0.5), nlp.hess_coord(x, obj_weight=0.5), **tol)
    # outputs written through views of larger buffers
    out = torch.full((nlp.meta.nnzh + 9,), float("nan"), dtype=torch.float64, device="cuda")
    nlp.hess_coord_(xv, out[5: 5 + nlp.meta.nnzh], obj_weight=0.5)
    assert torch.allclose(out[5: 5 + nlp.meta.nnzh], nlp.hess_coord(x, obj_weight=0.5), **tol)
    assert torch.isnan(out[:5]).all() and torch.isnan(out[5 + nlp.meta.nnzh:]).all()
    if m:
        assert torch.allclose(nlp.cons(xv), nlp.cons(x), **tol)
        assert torch.allclose(nlp.jac_coord(xv), nlp.jac_coord(x), **tol)
        assert torch.allclose(nlp.jprod(xv, vv), nlp.jprod(x, v), **tol)
        assert torch.allclose(nlp.jtprod(xv, yv), nlp.jtprod(x, y), **tol)
        assert torch.allclose(nlp.hess_coord(xv, yv, obj_weight=0.5), nlp.hess_coord(x, y, obj_weight=0.5), **tol)
        assert torch.allclose(nlp.hprod(xv, yv, vv, obj_weight=0.5), nlp.hprod(x, y, v, obj_weight=0.5), **tol)


def test_sis_residual_at_exact_solution_on_the_gpu(pkg, cuda_lib):
    """sis_test (test/problems/sis.jl:78-93) through the CUDA path: the weak residual of the exact ODE solution
    is below 1e-15 at n = 1e5 (the reference accepts the first n = 10^k, k <= 6, that gets there)."""
    from pdenlp_b200.model import GridapPDENLPModel
    from test_oracle import _sis_exact
    n = 10 ** 5
    nlp = GridapPDENLPModel(pkg.sis(n), device="cuda:0")
    x = _d(_sis_exact(n))
    assert float(nlp.cons(x).abs().max()) <= 1e-15
    assert nlp.obj(x) == 0.0 and float(nlp.grad(x).abs().max()) == 0.0
    nlp.close()


def test_generic_wrappers(prob):
    """objgrad, objcons, jac, hess: the NLPModels compositions of the callbacks (Appendix A.1)."""
    name, spec, nlp = prob
    n, m = nlp.meta.nvar, nlp.meta.ncon
    x = _d(_points(n)[2]); y = _d(np.random.default_rng(13).uniform(-1, 1, m))
    f, g = nlp.objgrad(x)
    assert f == nlp.obj(x) and torch.allclose(g, nlp.grad(x), rtol=1e-13, atol=1e-15)
    f2, c = nlp.objcons(x)
    assert f2 == f and c.numel() == m and (m == 0 or torch.allclose(c, nlp.cons(x), rtol=1e-13, atol=1e-15))
    if getattr(spec, "affine", False):
        return
    colptr, rowval, nzval = nlp.hess(x, y if m else None, obj_weight=0.7)
    H = sp.csc_matrix((_h(nzval), _h(rowval) - 1, _h(colptr) - 1), shape=(n, n)).toarray()
    r, c_ = nlp.hess_structure()
    vals = nlp.hess_coord(x, y, obj_weight=0.7) if m else nlp.hess_coord(x, obj_weight=0.7)
    Href = sp.coo_matrix((_h(vals), (_h(r) - 1, _h(c_) - 1)), shape=(n, n)).toarray()
    assert np.allclose(H, Href, rtol=1e-13, atol=1e-15) and np.all(np.triu(H, 1) == 0)
    if m:
        colptr, rowval, nzval = nlp.jac(x)
        J = sp.csc_matrix((_h(nzval), _h(rowval) - 1, _h(colptr) - 1), shape=(m, n)).toarray()
        r, c_ = nlp.jac_structure()
        Jref = sp.coo_matrix((_h(nlp.jac_coord(x)), (_h(r) - 1, _h(c_) - 1)), shape=(m, n)).toarray()
        assert np.allclose(J, Jref, rtol=1e-13, atol=1e-15)


def test_host_memspace_on_every_problem(prob):
    """PDENLP_MEM_HOST (numpy in / out, library-side staging, host zero fill of structurally zero blocks) gives the
    device-memspace results on every problem."""
    from pdenlp_b200.model import GridapPDENLPModel
    name, spec, nlp = prob
    if getattr(spec, "affine", False):
        pytest.skip("the linear-API model keeps its vectors on the device")
    nh = GridapPDENLPModel(spec, device="cuda:0", memspace="host")
    n, m = nlp.meta.nvar, nlp.meta.ncon
    rng = np.random.default_rng(21)
    x = rng.uniform(-1, 1, n); y = rng.uniform(-1, 1, m); v = rng.uniform(-1, 1, n)
    tol = dict(rtol=1e-12, atol=1e-14)
    assert np.isclose(nh.obj(x), nlp.obj(_d(x)), rtol=1e-13, atol=1e-15)
    assert np.allclose(nh.grad(x), _h(nlp.grad(_d(x))), **tol)
    fh, gh = nh.objgrad(x)
    assert np.isclose(fh, nlp.obj(_d(x)), rtol=1e-13, atol=1e-15) and np.allclose(gh, _h(nlp.grad(_d(x))), **tol)
    fh, ch = nh.objcons(x)
    assert np.isclose(fh, nlp.obj(_d(x)), rtol=1e-13, atol=1e-15) and (m == 0 or np.allclose(ch, _h(nlp.cons(_d(x))), **tol))
    assert np.allclose(nh.hess_coord(x, obj_weight=0.3), _h(nlp.hess_coord(_d(x), obj_weight=0.3)), **tol)
    assert np.allclose(nh.hprod(x, v, obj_weight=0.3), _h(nlp.hprod(_d(x), _d(v), obj_weight=0.3)), **tol)
    r, c = nh.hess_structure(); rd, cd = nlp.hess_structure()
    assert np.array_equal(r, _h(rd)) and np.array_equal(c, _h(cd))
    if m:
        assert np.allclose(nh.cons(x), _h(nlp.cons(_d(x))), **tol)
        assert np.allclose(nh.jac_coord(x), _h(nlp.jac_coord(_d(x))), **tol)
        assert np.allclose(nh.hess_coord(x, y, obj_weight=0.3), _h(nlp.hess_coord(_d(x), _d(y), obj_weight=0.3)), **tol)
        assert np.allclose(nh.jprod(x, v), _h(nlp.jprod(_d(x), _d(v))), **tol)
        assert np.allclose(nh.jtprod(x, y), _h(nlp.jtprod(_d(x), _d(y))), **tol)
        assert np.allclose(nh.hprod(x, y, v, obj_weight=0.3), _h(nlp.hprod(_d(x), _d(y), _d(v), obj_weight=0.3)), **tol)
        r, c = nh.jac_structure(); rd, cd = nlp.jac_structure()
        assert np.array_equal(r, _h(rd)) and np.array_equal(c, _h(cd))
    nh.close()
