"""SURVEY §8(f) rows 1 and 3 on the GPU: the linear-API path of an AffineFEOperator model
(cons_lin!, jac_lin_coord!, jac_lin_structure!, jprod_lin!, jtprod_lin!;
/root/reference/src/GridapPDENLPModel.jl:460-471,489-503,521-535,597-608) and the device
COO -> CSC compaction behind it (= NLPModels jac()/hess()).

Oracle: the assembled matrix findnz(get_matrix(op)) is restated with scipy from the oracle's COO
stream: duplicates summed, CSC order with rows ascending inside a column, stored zeros kept
(Julia's sparse(I,J,V) keeps numerical zeros)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


from parity import is_close as _close  # noqa: E402  (element-wise + segment-relative 1e-12, tests/parity.py)


def _findnz(rows, cols, vals, shape):
    """findnz(sparse(rows, cols, vals, m, n)): CSC order, duplicates summed in input order, zeros kept."""
    order = np.lexsort((rows, cols))  # stable: by col, then row, ties in input (= cell) order
    r, c, v = rows[order], cols[order], vals[order]
    head = np.ones(r.size, dtype=bool)
    head[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    idx = np.cumsum(head) - 1
    out = np.zeros(idx[-1] + 1)
    np.add.at(out, idx, v)
    return r[head], c[head], out


@pytest.mark.parametrize("n", [3, 8, 37])
def test_affine_membrane_linear_api(n, pkg, oracle_mod, cuda_lib):
    from pdenlp_b200.model import GridapPDENLPModel
    spec = pkg.controlelasticmembrane(n, affine=True)
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    orc = oracle_mod.Oracle(pkg.controlelasticmembrane(n))
    nvar, ncon = orc.nvar, orc.ncon
    rng = np.random.default_rng(1998)
    x = rng.uniform(-1, 1, nvar); lam = rng.uniform(-1, 1, ncon); v = rng.uniform(-1, 1, nvar)
    jr, jc = orc.jac_structure()
    Ar, Ac, Av = _findnz(jr, jc, orc.jac_coord(x), (ncon, nvar))
    A = sp.csc_matrix((Av, (Ar - 1, Ac - 1)), shape=(ncon, nvar))
    b = -orc.cons(np.zeros(nvar))
    # meta (src/additional_constructors.jl:213-242)
    m = nlp.meta
    assert (m.nlin, m.nnln, m.lin_nnzj, m.nln_nnzj, m.nnzj) == (ncon, 0, Ar.size, 0, Ar.size)
    assert m.nnzh == orc.nnzh_obj
    if n == 3:
        assert m.nnzh == 455  # the CI membrane problem
    # structure: bit-exact findnz order
    r, c = nlp.jac_structure()
    assert np.array_equal(r.cpu().numpy(), Ar) and np.array_equal(c.cpu().numpy(), Ac)
    # values
    assert _close(nlp.jac_coord(_d(x)).cpu().numpy(), Av)
    assert _close(nlp.cons(_d(x)).cpu().numpy(), A @ x - b)
    assert _close(nlp.jprod(_d(x), _d(v)).cpu().numpy(), A @ v)
    assert _close(nlp.jtprod(_d(x), _d(lam)).cpu().numpy(), A.T @ lam)
    # Hessian: objective part only, with or without multipliers
    ho = orc.hess_coord(x, None, 0.37)[: orc.nnzh_obj]
    assert _close(nlp.hess_coord(_d(x), obj_weight=0.37).cpu().numpy(), ho)
    assert _close(nlp.hess_coord(_d(x), _d(lam), obj_weight=0.37).cpu().numpy(), ho)
    hr, hc = nlp.hess_structure()
    ro, co = orc.hess_structure()
    assert np.array_equal(hr.cpu().numpy(), ro[: orc.nnzh_obj]) and np.array_equal(hc.cpu().numpy(), co[: orc.nnzh_obj])
    assert _close(nlp.hprod(_d(x), _d(lam), _d(v), obj_weight=0.37).cpu().numpy(), orc.hprod(x, v, None, 0.37))
    # counters: the _lin family
    c_ = nlp.counters
    assert (c_.neval_cons, c_.neval_cons_lin, c_.neval_cons_nln) == (1, 1, 0)
    assert (c_.neval_jac, c_.neval_jac_lin, c_.neval_jprod_lin, c_.neval_jtprod_lin) == (1, 1, 1, 1)
    # dimension errors
    from pdenlp_b200.model import DimensionError
    with pytest.raises(DimensionError):
        nlp.jac_coord_(_d(x), torch.empty(Ar.size + 1, dtype=torch.float64, device="cuda"))
    nlp.close()


@pytest.mark.parametrize("case", ["pb", "burger_nl", "membrane"])
def test_jac_and_hess_csc(case, pkg, oracle_mod, cuda_lib):
    """NLPModels jac(nlp,x) / hess(nlp,x,y) = sparse(rows, cols, vals): device compaction vs scipy."""
    from pdenlp_b200.model import GridapPDENLPModel
    spec = {"pb": lambda: pkg.poissonboltzman2d(12), "burger_nl": lambda: pkg.burger1d_param(50),
            "membrane": lambda: pkg.controlelasticmembrane(9)}[case]()
    nlp = GridapPDENLPModel(spec, device="cuda:0")
    orc = oracle_mod.Oracle(spec)
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, orc.nvar); lam = rng.uniform(-1, 1, orc.ncon)
    for kind in ("jac", "hess"):
        if kind == "jac":
            ptr, rows, vals = nlp.jac_csc(_d(x))
            r, c = orc.jac_structure(); v = orc.jac_coord(x); shape = (orc.ncon, orc.nvar)
        else:
            ptr, rows, vals = nlp.hess_csc(_d(x), _d(lam), obj_weight=0.37)
            r, c = orc.hess_structure(); v = orc.hess_coord(x, lam, 0.37); shape = (orc.nvar, orc.nvar)
        Rr, Rc, Rv = _findnz(r, c, v, shape)
        assert np.array_equal(rows.cpu().numpy(), Rr)
        ref = sp.csc_matrix((Rv, (Rr - 1, Rc - 1)), shape=shape)
        ref.sort_indices()
        colptr = np.concatenate([[0], np.cumsum(np.bincount(Rc - 1, minlength=shape[1]))]) + 1
        assert np.array_equal(ptr.cpu().numpy(), colptr)
        assert _close(vals.cpu().numpy(), Rv)
    nlp.close()


def test_compaction_csr_and_empty(cuda_lib):
    from pdenlp_b200 import capi
    rows = np.array([3, 1, 3, 2, 1, 3], dtype=np.int64); cols = np.array([2, 1, 2, 2, 1, 1], dtype=np.int64)
    vals = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], dtype=torch.float64, device="cuda")
    for order, expect in ((capi.ORDER_CSC, ([1, 3, 2, 3], [1, 1, 2, 2], [7.0, 6.0, 4.0, 4.0])),
                          (capi.ORDER_CSR, ([1, 2, 3, 3], [1, 2, 1, 2], [7.0, 4.0, 6.0, 4.0]))):
        plan = capi.Compaction(rows, cols, 3, 2, order)
        assert plan.nnz == 4
        r = torch.empty(4, dtype=torch.int64, device="cuda"); c = torch.empty(4, dtype=torch.int64, device="cuda")
        plan.structure(r.data_ptr(), c.data_ptr(), None, on_device=True)
        out = torch.empty(4, dtype=torch.float64, device="cuda")
        plan.apply(vals.data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        assert r.tolist() == expect[0] and c.tolist() == expect[1] and out.tolist() == expect[2]
        hr = np.empty(4, dtype=np.int64); hp = np.empty(3 if order == capi.ORDER_CSC else 4, dtype=np.int64)
        plan.structure(hr.ctypes.data, None, hp.ctypes.data, on_device=False)
        assert hr.tolist() == expect[0]
        assert hp.tolist() == ([1, 3, 5] if order == capi.ORDER_CSC else [1, 2, 3, 5])
        plan.close()
