"""Randomised invariants of the host-side layers (numbering, flattening, slab partition) against the oracle."""
import numpy as np
import pytest

from pdenlp_b200 import flatten as F
from pdenlp_b200.cartesian import CartesianDiscreteModel, LagrangianFESpace
from pdenlp_b200.sharded import SlabLayout, partition_cells


@pytest.mark.parametrize("seed", range(12))
def test_numbering_is_a_bijection_and_signs_partition_the_dofs(seed):
    rng = np.random.default_rng(seed)
    dim = int(rng.integers(1, 3))
    n = tuple(int(v) for v in rng.integers(1, 6, dim))
    order = int(rng.integers(1, 3))
    dom = sum(((0.0, float(k + 1)) for k in range(dim)), ())
    model = CartesianDiscreteModel(dom, n if dim > 1 else n[0])
    nent = 3 ** dim - 1
    tags = sorted(set(int(t) for t in rng.integers(1, nent + 1, int(rng.integers(0, nent)))))
    dirichlet = "boundary" if rng.random() < 0.3 else (tags or None)
    V = LagrangianFESpace(model, order, "H1", dirichlet, 0.5)
    ids = V.cell_dof_ids
    free = np.unique(ids[ids > 0]); diri = np.unique(-ids[ids < 0])
    assert np.array_equal(free, np.arange(1, V.nfree + 1)) and np.array_equal(diri, np.arange(1, V.ndirichlet + 1))
    nodes = int(np.prod([order * k + 1 for k in n]))
    assert V.nfree + V.ndirichlet == nodes and V.dirichlet_vals.shape == (V.ndirichlet,)
    assert all(len(set(row)) == len(row) for row in ids.tolist())  # no dof repeated inside a cell
    if dirichlet == "boundary":
        assert V.nfree == int(np.prod([max(order * k - 1, 0) for k in n]))
    L = LagrangianFESpace(model, order, "L2")
    assert L.nfree == ids.shape[0] * ids.shape[1] and np.array_equal(np.sort(L.cell_dof_ids.ravel()), np.arange(1, L.nfree + 1))


@pytest.mark.parametrize("seed", range(8))
def test_structure_counts_and_oracle_agreement_on_random_sizes(seed, pkg, oracle_mod):
    rng = np.random.default_rng(100 + seed)
    makers = [lambda n: pkg.controlelasticmembrane(n), lambda n: pkg.poissonboltzman2d(n, "H1" if n % 2 else "L2"),
              lambda n: pkg.burger1d_param(n + 2), lambda n: pkg.poissonmixed(n), lambda n: pkg.poissonmixed(n, 2, True),
              lambda n: pkg.dynamicsir(n + 1), lambda n: pkg.cellincrease(n + 1), lambda n: pkg.poissonparam(n + 1)]
    spec = makers[seed % len(makers)](int(rng.integers(1, 7)))
    fm = F.flatten(spec)
    o = oracle_mod.Oracle(spec)
    r, c = F.jac_structure(fm); ro, co = o.jac_structure()
    assert np.array_equal(r, ro) and np.array_equal(c, co)
    r, c, nobj = F.hess_structure(fm); ro, co = o.hess_structure()
    assert np.array_equal(r, ro) and np.array_equal(c, co) and nobj == o.nnzh_obj
    assert (r >= c).all() and r.min() >= 1 and r.max() <= o.nvar
    # closed-form per-cell counts used by the kernels: fy^2, fy*fu, fy(fy+1)/2 + fu*fy + fu(fu+1)/2
    fy = (fm.cell_dofs_y > 0).sum(axis=1); fu = (fm.cell_dofs_u > 0).sum(axis=1)
    assert o.nnzj_y == int((fy * fy).sum()) and o.nnzj_u == int((fy * fu).sum())
    assert o.nnzh_fe == int((fy * (fy + 1) // 2 + fu * fy + fu * (fu + 1) // 2).sum())


@pytest.mark.parametrize("ncells,world", [(1, 1), (31, 2), (32, 2), (33, 2), (1000, 3), (4194304, 8), (10, 8), (97, 5)])
def test_partition_cells_covers_all_cells_on_tile_boundaries(ncells, world):
    parts = partition_cells(ncells, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == ncells
    for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(a % 32 == 0 for a, _ in parts)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_layout_interface_sets(world, pkg):
    spec = pkg.controlelasticmembrane(12)
    lay = SlabLayout(spec, world)
    ny, nu = spec.Ypde.nfree, spec.Ycon.nfree
    assert lay.touch_y.shape == (world, ny) and lay.touch_y.any(axis=0).all() and lay.touch_u.any(axis=0).all()
    shared_y = np.nonzero(lay.touch_y.sum(axis=0) > 1)[0]
    assert np.array_equal(np.sort(lay.iface_con), shared_y)
    assert lay.iface_var.size == shared_y.size + int((lay.touch_u.sum(axis=0) > 1).sum())
    # a dof is shared by at most two consecutive slabs (cells are x-fastest rows)
    assert lay.touch_y.sum(axis=0).max() <= 2 and lay.touch_u.sum(axis=0).max() <= 2
