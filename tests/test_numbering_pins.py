"""The reference's own pins for the dof numbering / layout, re-expressed.

/root/reference/test/unit-test.jl:755-879 builds a 2x2 mesh on (-1,1)^2 with a
Q2 Dirichlet-boundary state space and a Q1 control space and checks
``bounds_functions_to_vectors`` against explicit arrays; those arrays pin Gridap's
free-dof order.  They hold for our host-side generator.
"""
import numpy as np
import pytest

from pdenlp_b200.cartesian import CartesianDiscreteModel, LagrangianFESpace
from pdenlp_b200 import flatten as F


def _spaces():
    model = CartesianDiscreteModel((-1, 1, -1, 1), (2, 2))
    Ypde = LagrangianFESpace(model, 2, "H1", "boundary", 0.0)
    Ycon = LagrangianFESpace(model, 1, "H1")
    mids = model.lo + (model.cell_multi_index() + 0.5) * model.h
    return model, Ypde, Ycon, mids


def test_unit_test_jl_828_q1_free_dof_order():
    model, Ypde, Ycon, mids = _spaces()
    umin = mids[:, 0] + mids[:, 1]
    assert Ycon.nfree == 9
    assert np.array_equal(Ycon.interpolate_cellwise_constant(umin), [-1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0])
    umax = mids[:, 0] ** 2 + mids[:, 1] ** 2
    assert np.array_equal(Ycon.interpolate_cellwise_constant(umax), 0.5 * np.ones(9))


def test_unit_test_jl_846_q2_dirichlet_free_dof_order():
    model, Ypde, Ycon, mids = _spaces()
    umin = mids[:, 0] + mids[:, 1]
    assert Ypde.nfree == 9 and Ypde.ndirichlet == 16
    assert np.array_equal(Ypde.interpolate_cellwise_constant(umin), [1.0, 0.0, 0.0, 1.0, 1.0, -1.0, 0.0, 0.0, 1.0])


def test_unit_test_jl_855_multifield_layout():
    """lvar of Example 2 = [state part ; control part]: x = [y ; u] with y first."""
    model, Ypde, Ycon, mids = _spaces()
    umin = mids[:, 0] + mids[:, 1]
    lvar = np.concatenate([Ypde.interpolate_cellwise_constant(umin), Ycon.interpolate_cellwise_constant(umin)])
    assert np.array_equal(lvar, [1.0, 0.0, 0.0, 1.0, 1.0, -1.0, 0.0, 0.0, 1.0, -1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0])


def test_survey_a4_cell1_ids_and_hessian_entries(pkg):
    model, Ypde, Ycon, _ = _spaces()
    assert Ypde.cell_dof_ids[0].tolist() == [-1, -2, -4, 1, -9, 2, -10, 3, 6]
    assert Ycon.cell_dof_ids[0].tolist() == [1, 2, 4, 5]
    spec = pkg.controlelasticmembrane(2)
    fm = F.flatten(spec, (0, 1))
    rows, cols, nobj = F.hess_structure(fm)
    got = list(zip(rows[:nobj].tolist(), cols[:nobj].tolist()))
    b11 = [(1, 1), (2, 1), (3, 1), (6, 1), (2, 2), (3, 2), (6, 2), (3, 3), (6, 3), (6, 6)]
    b21 = [(r, c) for c in (1, 2, 3, 6) for r in (10, 11, 13, 14)]
    b22 = [(10, 10), (11, 10), (13, 10), (14, 10), (11, 11), (13, 11), (14, 11), (13, 13), (14, 13), (14, 14)]
    assert got == b11 + b21 + b22 and nobj == 36


def test_survey_size_table_small_and_config1(pkg, oracle_mod):
    # README res-variant at n = 3: nvar 41, ncon 25, nnzj 289+196, nnzh 2*455
    o = oracle_mod.Oracle(pkg.controlelasticmembrane(3))
    assert (o.nvar, o.ncon, o.nnzj_y, o.nnzj_u, o.nnzh) == (41, 25, 289, 196, 910)
    # config 1 (n = 100)
    spec = pkg.controlelasticmembrane(100)
    assert (spec.Ypde.nfree, spec.Ycon.nfree) == (39601, 10201)
    fm = F.flatten(spec, with_coef=False)
    r, c = F.jac_structure(fm)
    assert r.size == 792100 + 355216
    r, c, nobj = F.hess_structure(fm)
    assert (nobj, r.size) == (895668, 1791336)


def test_hessian_mask_is_on_global_ids(pkg):
    """for an interior Q2 cell the kept pair (top edge, left edge) sits in the UPPER
    local triangle because global edge ids are first-encounter (bottom < left < top < right)."""
    spec = pkg.controlelasticmembrane(4)
    ids = spec.Ypde.cell_dof_ids[5]  # interior cell (1,1)
    assert (ids > 0).all()
    top, left = ids[5], ids[6]
    assert left < top
    fm = F.flatten(spec, (5, 6))
    rows, cols, nobj = F.hess_structure(fm)
    pairs = set(zip(rows[:nobj].tolist(), cols[:nobj].tolist()))
    assert (int(top), int(left)) in pairs and (int(left), int(top)) not in pairs


def test_burger_dirichlet_tags(pkg):
    spec = pkg.burger1d(8)
    Y = spec.Ypde
    assert Y.nfree == 7 and Y.ndirichlet == 2
    assert Y.cell_dof_ids[0].tolist() == [-1, 1] and Y.cell_dof_ids[-1].tolist() == [7, -2]
    assert Y.dirichlet_vals.tolist() == [0.0, -1.0]  # uD0 = 0 on "diri0", uD1 = -1 on "diri1"


def test_l2_control_numbering(pkg):
    spec = pkg.poissonboltzman2d(3)
    assert spec.Ycon.nfree == 9 * 4
    assert spec.Ycon.cell_dof_ids[2].tolist() == [9, 10, 11, 12]


def test_bounds_functions_to_vectors_reference_vectors(pkg):
    """test/unit-test.jl:755-879 "Functions used to create bounds in constructors": the reference's own
    expected lvar / uvar (cell-midpoint sampling + last-cell-wins interpolation on the Gridap numbering)."""
    from pdenlp_b200.bounds import bounds_functions_to_vectors
    from pdenlp_b200.cartesian import CartesianDiscreteModel, LagrangianFESpace
    from pdenlp_b200.model_errors import DimensionError
    model = CartesianDiscreteModel((-1, 1, -1, 1), (2, 2))
    Ypde = LagrangianFESpace(model, 2, "H1", "boundary", 0.0)
    Ycon = LagrangianFESpace(model, 1, "H1")
    assert Ypde.nfree == 9 and Ycon.nfree == 9
    inf9 = np.full(9, np.inf)
    umin = lambda X: X[:, 0] + X[:, 1]
    umax = lambda X: X[:, 0] ** 2 + X[:, 1] ** 2
    exp_y = np.array([1.0, 0.0, 0.0, 1.0, 1.0, -1.0, 0.0, 0.0, 1.0])
    exp_u = np.array([-1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0])
    for T in (np.float16, np.float32, np.float64):
        # Example 0 / 0bis: vectors
        l, u = bounds_functions_to_vectors(Ycon, Ypde, model, -inf9, inf9, -inf9, inf9, dtype=T)
        assert np.array_equal(l, -np.inf * np.ones(18)) and np.array_equal(u, np.inf * np.ones(18)) and l.dtype == T
        l, u = bounds_functions_to_vectors(None, Ypde, model, -inf9, inf9, [], [], dtype=T)
        assert np.array_equal(l, -inf9) and np.array_equal(u, inf9) and u.dtype == T
        # Example 1: state bounds as vectors, control bounds as functions
        l, u = bounds_functions_to_vectors(Ycon, Ypde, model, -inf9, inf9, umin, umax, dtype=T)
        assert np.array_equal(l, np.concatenate([-inf9, exp_u])) and np.array_equal(u, np.concatenate([inf9, 0.5 * np.ones(9)]))
        # Example 1bis: VoidFESpace control, state bounds as functions
        l, u = bounds_functions_to_vectors(None, Ypde, model, umin, umax, None, None, dtype=T)
        assert np.array_equal(l, exp_y) and np.array_equal(u, 0.5 * np.ones(9)) and l.dtype == T
        # Example 2: everything as (vector-valued) functions
        vmin = lambda X: umin(X)[:, None]
        vmax = lambda X: umax(X)[:, None]
        l, u = bounds_functions_to_vectors(Ycon, Ypde, model, vmin, vmax, vmin, vmax, dtype=T)
        assert np.array_equal(l, np.concatenate([exp_y, exp_u])) and np.array_equal(u, 0.5 * np.ones(18))
    with pytest.raises(DimensionError):
        bounds_functions_to_vectors(Ycon, Ypde, model, -np.ones(8), inf9, -inf9, inf9)
    # multi-field spaces: one column per field
    m1 = CartesianDiscreteModel((0, 1), 4)
    fields = [LagrangianFESpace(m1, 1, "H1", [1], 1.0), LagrangianFESpace(m1, 1, "H1", [1], 2.0)]
    lo = lambda X: np.stack([X[:, 0], -X[:, 0]], axis=1)
    hi = lambda X: np.stack([X[:, 0] + 1, 1 - X[:, 0]], axis=1)
    l, u = bounds_functions_to_vectors(None, fields, m1, lo, hi)
    mids = np.array([0.125, 0.375, 0.625, 0.875])  # free dofs = nodes 2..5, last cell wins
    assert np.allclose(l, np.concatenate([mids[[1, 2, 3, 3]], -mids[[1, 2, 3, 3]]]))
    assert np.allclose(u, np.concatenate([mids[[1, 2, 3, 3]] + 1, 1 - mids[[1, 2, 3, 3]]]))
