"""Error convention of the C ABI on a real device: a negative PDENLP_E* code plus a message from
pdenlp_last_error(), never a crash or a silent fallback (include/pdenlp_cuda.h:14-33)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fm(pkg, cuda_lib):
    from pdenlp_b200 import flatten
    return flatten.flatten(pkg.controlelasticmembrane(4))


def _raises(capi, fm, code, fragment):
    with pytest.raises(capi.PdenlpError) as ei:
        capi.Handle(fm)
    assert ei.value.code == code and fragment in str(ei.value), str(ei.value)


def test_dof_ids_are_validated(fm, cuda_lib):
    from pdenlp_b200 import capi
    bad = copy.copy(fm)
    bad.cell_dofs_y = fm.cell_dofs_y.copy(); bad.cell_dofs_y[3, 2] = fm.nfree_y + 1
    _raises(capi, bad, capi.EINVAL, "state dof id out of range")
    bad = copy.copy(fm)
    bad.cell_dofs_u = fm.cell_dofs_u.copy(); bad.cell_dofs_u[0, 0] = 0
    _raises(capi, bad, capi.EINVAL, "control dof id out of range")
    bad = copy.copy(fm)
    bad.cell_dofs_y = fm.cell_dofs_y.copy(); bad.cell_dofs_y[5, 1] = -(fm.dir_y.size + 1)
    _raises(capi, bad, capi.EINVAL, "state dof id out of range")


def test_repeated_dof_in_a_cell_is_rejected(fm, cuda_lib):
    from pdenlp_b200 import capi
    bad = copy.copy(fm)
    ids = fm.cell_dofs_y.copy()
    free = np.argwhere(ids > 0)
    c = free[0][0]
    cols = np.nonzero(ids[c] > 0)[0]
    if cols.size < 2:
        c = [r for r in range(ids.shape[0]) if (ids[r] > 0).sum() >= 2][0]
        cols = np.nonzero(ids[c] > 0)[0]
    ids[c, cols[1]] = ids[c, cols[0]]
    bad.cell_dofs_y = ids
    _raises(capi, bad, capi.EINVAL, "repeats a dof id")


def test_unsupported_element_and_abi(fm, cuda_lib, monkeypatch):
    from pdenlp_b200 import capi
    bad = copy.copy(fm)
    bad.nq = 3  # (integrand, ny, nu, nq) outside the closed set of instantiated kernels
    bad.phi_y = np.tile(fm.phi_y, (3, 1)); bad.grad_y = np.tile(fm.grad_y, (3, 1, 1))
    bad.phi_u = np.tile(fm.phi_u, (3, 1)); bad.wdet = np.tile(fm.wdet, 3)
    _raises(capi, bad, capi.EUNSUPPORTED, "closed set")
    monkeypatch.setattr(capi, "ABI_VERSION", capi.ABI_VERSION + 1)
    _raises(capi, fm, capi.EINVAL, "abi_version")


def test_callbacks_that_need_coefficients_say_so(fm, cuda_lib):
    import torch
    from pdenlp_b200 import capi
    h = capi.Handle(fm, with_coef=False)
    x = torch.zeros(h.info.nvar, dtype=torch.float64, device="cuda")
    g = torch.zeros(h.info.nvar, dtype=torch.float64, device="cuda")
    with pytest.raises(capi.PdenlpError) as ei:
        h.obj(x.data_ptr())
    assert ei.value.code == capi.EINVAL and "coefficient" in str(ei.value)
    with pytest.raises(capi.PdenlpError):
        h.grad(x.data_ptr(), g.data_ptr())
    # the coordinate callbacks of this integrand do not read the coefficient fields
    v = torch.empty(h.info.nnzj, dtype=torch.float64, device="cuda")
    h.jac_coord(x.data_ptr(), v.data_ptr())
    torch.cuda.synchronize()
    assert torch.isfinite(v).all()
    h.close()


@pytest.mark.parametrize("degree", [1, 2])
def test_wrong_shortcut_trait_is_refused_by_the_create_time_self_check(pkg, cuda_lib, degree):
    """The shortcut traits (RES_FE_HESSIAN_ZERO: the constraint FE Hessian block is a zero fill;
    FE_DERIVS_POINT_INDEPENDENT: one quadrature point + pre-integrated reference tensors) are hand-set per integrand.
    pdenlp_create runs the shortcut AND the generic kernels on a sample of the model's cells and must refuse a
    model whose integrand declares them wrongly: the test-only integrand PDENLP_SELFTEST_WRONG_TRAIT is
    Poisson-Boltzmann (sinh term: neither trait holds) with both traits set."""
    from pdenlp_b200 import capi, flatten
    spec = pkg.poissonboltzman2d(12, "H1", degree=degree)
    good = flatten.flatten(spec)
    h = capi.Handle(good)          # the honest integrand passes (no shortcut applies: nothing to compare)
    assert h.selfcheck_error() == 0.0
    h.close()
    import copy as _copy
    bad = _copy.copy(good)
    bad.spec = _copy.copy(spec); bad.spec.integrand = 1000
    with pytest.raises(capi.PdenlpError) as ei:
        capi.Handle(bad)
    assert ei.value.code == capi.EINVAL and "self-check failed" in str(ei.value), str(ei.value)


def test_self_check_passes_and_reports_for_models_with_shortcuts(pkg, cuda_lib):
    """Models whose integrands really have the traits pass, with a disagreement at round-off level."""
    from pdenlp_b200 import capi, flatten
    for spec in (pkg.controlelasticmembrane(10), pkg.controlelasticmembrane(9, degree=2), pkg.poissonmixed(6, 3, True),
                 pkg.poissonmixed(7), pkg.burger1d(50, degree=2), pkg.poissonparam(8)):
        h = capi.Handle(flatten.flatten(spec))
        assert h.selfcheck_error() <= 1e-13, (spec.name, h.selfcheck_error())
        h.close()
