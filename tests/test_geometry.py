"""Per-cell geometry (pdenlp_desc ABI v3: cell_jinvT / cell_detj; Gridap's `CartesianDiscreteModel(...; map = f)`,
and through it any triangulation the reference constructor accepts, src/additional_constructors.jl:130-175).

CPU part: the oracle's mapped-mesh geometry (its own C++ code: vertex coordinates -> Q1 cell maps) against (i) the plain
Cartesian code on the same cells, (ii) finite differences on a genuinely warped mesh, (iii) the independent numpy
geometry of the flattening layer.  GPU part: the CUDA per-cell-geometry kernels against the oracle on warped meshes,
and against the uniform fast path on meshes that are in fact uniform."""
import numpy as np
import pytest

from parity import assert_parity


def warp2(X):
    """Smooth interior warp of (-1,1)^2 onto itself (boundary fixed): cells become general, non-affine quads."""
    x, y = X[:, 0], X[:, 1]
    return np.stack([x + 0.11 * np.sin(np.pi * x) * np.cos(0.5 * np.pi * y), y - 0.09 * np.sin(np.pi * y) * np.cos(0.5 * np.pi * x)], axis=1)


def warp01(X):
    """The same kind of warp on (0,1)^dim."""
    Y = X.copy()
    s = np.prod(np.sin(np.pi * X), axis=1)
    for d in range(X.shape[1]):
        Y[:, d] = X[:, d] + (0.05 + 0.02 * d) * s
    return Y


def stretch1(X):
    return X ** 1.7 * 0.6 + 0.4 * X   # monotone map of (0,1) onto itself: graded 1-D mesh


def shear2(X):
    """Affine map (shear + anisotropic scaling + shift): every cell is the same parallelogram."""
    A = np.array([[1.3, 0.4], [-0.2, 0.8]])
    return X @ A.T + np.array([0.3, -0.1])


def _vecs(o, spec, seed=11):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, o.nvar)
    if spec.nparam:
        x[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
    return x, rng.uniform(-1, 1, o.ncon), rng.uniform(-1, 1, o.nvar)


@pytest.mark.parametrize("case", ["membrane", "pb_deg2", "kappa3d", "burger_nl"])
def test_identity_map_reproduces_the_plain_cartesian_oracle(case, pkg, oracle_mod):
    """map = identity runs the oracle's mapped-geometry code (vertex coordinates, Q1 cell maps, J^-T, det J) on cells
    that are the plain Cartesian ones: every output must agree with the plain code to round-off."""
    ident = lambda X: X.copy()
    mk = {"membrane": lambda m: pkg.controlelasticmembrane(7, mesh_map=m), "pb_deg2": lambda m: pkg.poissonboltzman2d(6, degree=2, mesh_map=m),
          "kappa3d": lambda m: pkg.poissonmixed(3, 3, True, mesh_map=m), "burger_nl": lambda m: pkg.burger1d_param(40, mesh_map=m)}[case]
    a, b = oracle_mod.Oracle(mk(None), nthreads=1), oracle_mod.Oracle(mk(ident), nthreads=1)
    x, lam, v = _vecs(a, a.spec)
    tol = dict(rtol=1e-13, atol=1e-14)
    assert np.isclose(a.obj(x), b.obj(x), **tol) and np.allclose(a.grad(x), b.grad(x), **tol) and np.allclose(a.cons(x), b.cons(x), **tol)
    assert np.allclose(a.jac_coord(x), b.jac_coord(x), **tol) and np.allclose(a.hess_coord(x, lam, 0.7), b.hess_coord(x, lam, 0.7), **tol)
    assert np.array_equal(a.jac_structure()[0], b.jac_structure()[0]) and np.array_equal(a.hess_structure()[1], b.hess_structure()[1])


@pytest.mark.parametrize("case", ["membrane_warp", "pb_warp_deg2", "kappa3d_warp", "burger_nl_graded"])
def test_oracle_derivatives_on_a_warped_mesh_pass_finite_difference_checks(case, pkg, oracle_mod):
    mk = {"membrane_warp": lambda: pkg.controlelasticmembrane(5, mesh_map=warp2), "pb_warp_deg2": lambda: pkg.poissonboltzman2d(5, "H1", degree=2, mesh_map=warp2),
          "kappa3d_warp": lambda: pkg.poissonmixed(3, 3, True, mesh_map=warp01), "burger_nl_graded": lambda: pkg.burger1d_param(30, mesh_map=stretch1)}[case]
    spec = mk()
    o = oracle_mod.Oracle(spec, nthreads=1)
    x, lam, v = _vecs(o, spec)
    h = 1e-6
    fd = (o.obj(x + h * v) - o.obj(x - h * v)) / (2 * h)
    assert np.isclose(fd, o.grad(x) @ v, rtol=1e-7, atol=1e-9)
    cfd = (o.cons(x + h * v) - o.cons(x - h * v)) / (2 * h)
    assert np.allclose(cfd, o.jprod(x, v), rtol=1e-6, atol=1e-8)
    lag = lambda z: o.grad(z) * 0.7 + o.jtprod(z, lam)
    hfd = (lag(x + h * v) - lag(x - h * v)) / (2 * h)
    assert np.allclose(hfd, o.hprod(x, v, lam, 0.7), rtol=1e-5, atol=1e-7)


def test_flatten_geometry_matches_the_mesh(pkg):
    """numpy geometry of the flattening layer: sum of w_q |det J| = area of the (boundary-preserving) warped
    domain; J^-T J^T = I; an affine map gives the same Jacobian in every cell and at every point."""
    from pdenlp_b200.flatten import flatten
    fm = flatten(pkg.controlelasticmembrane(12, degree=3, mesh_map=warp2))
    assert fm.geom_nq == fm.nq == 4 and fm.cell_jinvT.shape == (144, 4, 2, 2)
    assert np.isclose((fm.cell_detj * fm.wdet[None, :]).sum(), 4.0, rtol=1e-3)   # Q1 cells approximate the warped cells
    m = pkg.controlelasticmembrane(12, degree=3, mesh_map=warp2).model
    J, det = m.cell_jacobians(np.arange(m.ncells), np.array([[0.3, 0.6]]))
    assert np.allclose(np.einsum("cqde,cqfe->cqdf", fm.cell_jinvT[:, :1] * 0 + np.swapaxes(np.linalg.inv(J), -1, -2), J), np.eye(2)[None, None], atol=1e-12)
    fa = flatten(pkg.controlelasticmembrane(6, mesh_map=shear2))
    A = np.array([[1.3, 0.4], [-0.2, 0.8]]) * (2.0 / 6)      # J = A * diag(h)
    assert np.allclose(fa.cell_jinvT, np.linalg.inv(A).T[None, None], rtol=1e-12) and np.allclose(fa.cell_detj, abs(np.linalg.det(A)))
    # a plain Cartesian mesh takes the uniform fast path unless asked otherwise
    fu = flatten(pkg.controlelasticmembrane(6))
    assert fu.cell_jinvT is None and fu.geom_nq == 0
    fp = flatten(pkg.controlelasticmembrane(6), geometry="affine")
    assert fp.geom_nq == 1 and np.allclose(fp.cell_jinvT, np.diag(1 / fu.spec.model.h)[None, None]) and np.allclose(fp.grad_y * 6 / 2, fu.grad_y)


# ------------------------------------------------------------------------------------------------- GPU
GPU_CASES = {
    "membrane_q2_warp": lambda pkg: pkg.controlelasticmembrane(9, mesh_map=warp2),
    "membrane_q2_warp_deg3": lambda pkg: pkg.controlelasticmembrane(8, degree=3, mesh_map=warp2),
    "membrane_shear": lambda pkg: pkg.controlelasticmembrane(10, mesh_map=shear2),
    "pb_l2_warp": lambda pkg: pkg.poissonboltzman2d(11, mesh_map=warp2),
    "pb_h1_warp_deg2": lambda pkg: pkg.poissonboltzman2d(9, "H1", degree=2, mesh_map=warp2),
    "pb_q2_warp_deg2": lambda pkg: pkg.poissonboltzman2d(7, state_order=2, degree=2, mesh_map=warp2),
    "kappa2d_warp": lambda pkg: pkg.poissonmixed(7, 2, True, mesh_map=warp01),
    "kappa3d_warp": lambda pkg: pkg.poissonmixed(4, 3, True, mesh_map=warp01),
    "membrane3d_warp_deg2": lambda pkg: pkg.controlelasticmembrane(4, state_order=1, dim=3, degree=2,
                                                                   mesh_map=lambda X: 2 * warp01(0.5 * (X + 1)) - 1),
    "burger_lin_graded_deg2": lambda pkg: pkg.burger1d(60, degree=2, mesh_map=stretch1),
    "burger_nl_graded": lambda pkg: pkg.burger1d_param(70, mesh_map=stretch1),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GPU_CASES))
def test_cuda_per_cell_geometry_matches_the_oracle(name, pkg, oracle_mod, cuda_lib):
    import torch
    from pdenlp_b200.model import GridapPDENLPModel
    spec = GPU_CASES[name](pkg)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    assert nlp.flat.geom_nq == nlp.flat.nq          # mapped mesh: Jacobians at every quadrature point
    o = oracle_mod.Oracle(spec)
    x, lam, v = _vecs(o, spec)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    h = lambda t: t.cpu().numpy()
    r, c = nlp.jac_structure(); ro, co = o.jac_structure()
    assert np.array_equal(h(r), ro) and np.array_equal(h(c), co)
    r, c = nlp.hess_structure(); ro, co = o.hess_structure()
    assert np.array_equal(h(r), ro) and np.array_equal(h(c), co)
    f, fo = nlp.obj(d(x)), o.obj(x)
    assert abs(f - fo) <= 1e-12 * abs(fo)
    assert_parity(h(nlp.grad(d(x))), o.grad(x), "grad")
    assert_parity(h(nlp.cons(d(x))), o.cons(x), "cons")
    assert_parity(h(nlp.jac_coord(d(x))), o.jac_coord(x), "jac_coord")
    no = o.nnzh_obj
    for ow in (1.0, 0.37):
        got, ref = h(nlp.hess_coord(d(x), d(lam), obj_weight=ow)), o.hess_coord(x, lam, ow)
        assert_parity(got[:no], ref[:no], "hess_coord obj"); assert_parity(got[no:], ref[no:], "hess_coord cons")
    assert_parity(h(nlp.hess_coord(d(x), obj_weight=0.37)), o.hess_coord(x, None, 0.37), "hess_coord(x)")
    assert_parity(h(nlp.jprod(d(x), d(v))), o.jprod(x, v), "jprod")
    assert_parity(h(nlp.jtprod(d(x), d(lam))), o.jtprod(x, lam), "jtprod")
    assert_parity(h(nlp.hprod(d(x), d(lam), d(v), obj_weight=0.37)), o.hprod(x, v, lam, 0.37), "hprod")
    nlp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("geometry", ["per_point", "affine"])
@pytest.mark.parametrize("name", ["membrane", "pb_deg2", "kappa3d", "burger_nl", "dynamicsir"])
def test_explicit_geometry_of_a_uniform_mesh_matches_the_fast_path(name, geometry, pkg, cuda_lib):
    """The same uniform mesh described through cell_jinvT / cell_detj (per point, or one Jacobian per cell) must
    reproduce the uniform fast path: structure bit-exact, values to round-off."""
    import torch
    from pdenlp_b200.model import GridapPDENLPModel
    spec = {"membrane": lambda: pkg.controlelasticmembrane(12), "pb_deg2": lambda: pkg.poissonboltzman2d(9, degree=2),
            "kappa3d": lambda: pkg.poissonmixed(4, 3, True), "burger_nl": lambda: pkg.burger1d_param(64),
            "dynamicsir": lambda: pkg.dynamicsir(50)}[name]()
    fast = GridapPDENLPModel(spec, device="cuda:0")
    geo = GridapPDENLPModel(spec, device="cuda:0", geometry=geometry)
    assert fast.flat.geom_nq == 0 and geo.flat.geom_nq == (1 if geometry == "affine" else geo.flat.nq)
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, fast.meta.nvar)
    if spec.nparam:
        x[: spec.nparam] = 1 + 0.1 * rng.uniform(-1, 1, spec.nparam)
    x, lam, v = (torch.from_numpy(a).cuda() for a in (x, rng.uniform(-1, 1, fast.meta.ncon), rng.uniform(-1, 1, fast.meta.nvar)))
    h = lambda t: t.cpu().numpy()
    assert torch.equal(fast.jac_structure()[0], geo.jac_structure()[0]) and torch.equal(fast.hess_structure()[1], geo.hess_structure()[1])
    assert abs(fast.obj(x) - geo.obj(x)) <= 1e-13 * abs(fast.obj(x))
    for a, b, what in ((geo.grad(x), fast.grad(x), "grad"), (geo.cons(x), fast.cons(x), "cons"),
                       (geo.jac_coord(x), fast.jac_coord(x), "jac"), (geo.hess_coord(x, lam, obj_weight=0.6), fast.hess_coord(x, lam, obj_weight=0.6), "hess"),
                       (geo.jprod(x, v), fast.jprod(x, v), "jprod"), (geo.jtprod(x, lam), fast.jtprod(x, lam), "jtprod"),
                       (geo.hprod(x, lam, v, obj_weight=0.6), fast.hprod(x, lam, v, obj_weight=0.6), "hprod")):
        assert_parity(h(a), h(b), what, rtol=1e-13)
    fast.close(); geo.close()
