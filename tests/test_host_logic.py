"""Host-side mirror of the reference interface, exercised without a GPU through a recording
stand-in for the C-ABI handle: @lencheck -> DimensionError before any work, Counters
bookkeeping (test/unit-test.jl:881-925), argument routing of the two hess_coord!/hprod! methods."""
import numpy as np
import pytest
import torch

from pdenlp_b200 import capi, model as M


class _Info:
    nvar, ncon, nnzj, nnzj_k, nnzj_y, nnzj_u = 41, 25, 485, 0, 289, 196
    nnzh, nnzh_obj, nnzh_k_obj, nnzh_fe, nnzh_k_con = 910, 455, 0, 455, 0


class FakeHandle:
    def __init__(self, fm, device=-1, memspace=0, with_coef=True, flags=0):
        self.info = _Info(); self.calls = []; self.memspace = memspace

    def set_stream(self, s): pass
    def close(self): pass
    def launch_count(self): return len(self.calls)
    def obj(self, x): self.calls.append(("obj", x)); return 1.5
    def grad(self, x, g): self.calls.append(("grad", x, g))
    def objgrad(self, x, g): self.calls.append(("objgrad", x, g)); return 1.5
    def objcons(self, x, c): self.calls.append(("objcons", x, c)); return 1.5
    def cons(self, x, c): self.calls.append(("cons", x, c))
    def jac_coord(self, x, v): self.calls.append(("jac_coord", x, v))
    def hess_coord(self, x, lam, ow, v): self.calls.append(("hess_coord", x, lam, ow, v))
    def jprod(self, x, v, o): self.calls.append(("jprod", x, v, o))
    def jtprod(self, x, v, o): self.calls.append(("jtprod", x, v, o))
    def hprod(self, x, lam, ow, v, o): self.calls.append(("hprod", x, lam, ow, v, o))
    def jac_structure(self, r, c): self.calls.append(("jac_structure", r, c))
    def hess_structure(self, r, c): self.calls.append(("hess_structure", r, c))


@pytest.fixture()
def nlp(monkeypatch, pkg):
    monkeypatch.setattr(capi, "Handle", FakeHandle)
    monkeypatch.setattr(M.GridapPDENLPModel, "_use_current_stream", lambda self: None)
    return M.GridapPDENLPModel(pkg.controlelasticmembrane(3), device="cuda:0", memspace="host")


def test_meta(nlp):
    assert (nlp.meta.nvar, nlp.meta.ncon, nlp.meta.nnzj, nlp.meta.nnzh) == (41, 25, 485, 910)
    assert nlp.meta.nlin == 0 and nlp.meta.nnln == 25 and nlp.meta.nln_nnzj == 485
    assert nlp.pdemeta.nnzh_obj == 455 and nlp.pdemeta.nparam == 0


def test_dimension_errors_before_any_work(nlp):
    x = np.zeros(41); bad = np.zeros(40)
    cases = [
        lambda: nlp.obj(bad), lambda: nlp.grad_(bad, np.zeros(41)), lambda: nlp.grad_(x, np.zeros(40)),
        lambda: nlp.cons_(x, np.zeros(24)), lambda: nlp.cons_(bad, np.zeros(25)),
        lambda: nlp.jac_coord_(x, np.zeros(484)), lambda: nlp.jac_coord_(bad, np.zeros(485)),
        lambda: nlp.hess_coord_(x, np.zeros(909)), lambda: nlp.hess_coord_(x, np.zeros(24), np.zeros(910)),
        lambda: nlp.hess_coord_(bad, np.zeros(25), np.zeros(910)),
        lambda: nlp.jprod_(x, np.zeros(40), np.zeros(25)), lambda: nlp.jprod_(x, x, np.zeros(26)),
        lambda: nlp.jtprod_(x, np.zeros(24), np.zeros(41)), lambda: nlp.jtprod_(x, np.zeros(25), np.zeros(40)),
        lambda: nlp.hprod_(x, np.zeros(40), np.zeros(41)), lambda: nlp.hprod_(x, np.zeros(24), x, np.zeros(41)),
        lambda: nlp.hess_structure_(np.zeros(909, np.int64), np.zeros(910, np.int64)),
        lambda: nlp.jac_structure_(np.zeros(485, np.int64), np.zeros(486, np.int64)),
    ]
    for f in cases:
        with pytest.raises(M.DimensionError):
            f()
    assert nlp.handle.calls == []  # nothing reached the library
    assert all(v == 0 for v in nlp.counters.as_dict().values())


def test_counters_match_reference_bookkeeping(nlp):
    x = np.zeros(41); lam = np.zeros(25)
    nlp.obj(x); nlp.grad(x); nlp.cons(x); nlp.jac_coord(x)
    nlp.hess_coord(x); nlp.hess_coord(x, lam)
    c = nlp.counters
    assert (c.neval_obj, c.neval_grad, c.neval_cons, c.neval_cons_nln, c.neval_jac, c.neval_jac_nln, c.neval_hess) == (1, 1, 1, 1, 1, 1, 2)
    # hprod!/jprod!/jtprod! bump their own counter; the nested hess/jac evaluation does not count
    nlp.hprod(x, x); nlp.hprod(x, lam, x); nlp.jprod(x, x); nlp.jtprod(x, lam)
    assert (c.neval_hprod, c.neval_hess, c.neval_jprod, c.neval_jprod_nln, c.neval_jtprod, c.neval_jtprod_nln) == (2, 2, 1, 1, 1, 1)
    assert c.neval_jac == 1 and c.neval_jac_nln == 1
    nlp.reset_()
    assert all(v == 0 for v in c.as_dict().values())


def test_method_routing(nlp):
    x = np.zeros(41); lam = np.ones(25); v = np.ones(41)
    nlp.hess_coord_(x, np.zeros(910), obj_weight=0.5)
    assert nlp.handle.calls[-1][0] == "hess_coord" and nlp.handle.calls[-1][2] is None and nlp.handle.calls[-1][3] == 0.5
    nlp.hess_coord_(x, lam, np.zeros(910))
    assert nlp.handle.calls[-1][2] is not None and nlp.handle.calls[-1][3] == 1.0
    nlp.hprod_(x, v, np.zeros(41), obj_weight=0.0)
    assert nlp.handle.calls[-1][0] == "hprod" and nlp.handle.calls[-1][2] is None
    nlp.hprod_(x, lam, v, np.zeros(41))
    assert nlp.handle.calls[-1][2] is not None
    prod, tprod = nlp.jac_op(x)
    prod(v); tprod(lam)
    assert [c[0] for c in nlp.handle.calls[-2:]] == ["jprod", "jtprod"]
    nlp.hess_op(x, lam)(v)
    assert nlp.handle.calls[-1][0] == "hprod"


def test_views_are_accepted(nlp):
    """view_subarray_nlp (test/runtests.jl:78): strided views are staged, results land in the view."""
    big = np.zeros(82)
    xv = big[::2]
    out = np.full(82, 7.0)
    gv = out[::2]
    nlp.grad_(xv, gv)
    assert nlp.handle.calls[-1][0] == "grad"
    t = torch.zeros(82, dtype=torch.float64)[::2]
    nlp.grad_(t, torch.zeros(41, dtype=torch.float64))


def test_constructor_bounds_and_x0(monkeypatch, pkg):
    """src/additional_constructors.jl:178-204: bounds on y / u as functions or vectors, kappa bounds in front."""
    monkeypatch.setattr(capi, "Handle", FakeHandle)
    monkeypatch.setattr(M.GridapPDENLPModel, "_use_current_stream", lambda self: None)
    spec = pkg.controlelasticmembrane(3)
    umin = lambda X: X[:, 0] + X[:, 1]
    umax = lambda X: X[:, 0] ** 2 + X[:, 1] ** 2 + 10.0
    nlp = M.GridapPDENLPModel(spec, device="cuda:0", memspace="host", lvaru=umin, uvaru=umax, x0=np.ones(41))
    ny = spec.Ypde.nfree
    assert np.all(np.isinf(nlp.meta.lvar[:ny])) and np.all(np.isinf(nlp.meta.uvar[:ny]))
    assert np.all(np.isfinite(nlp.meta.lvar[ny:])) and np.all(nlp.meta.lvar[ny:] <= nlp.meta.uvar[ny:])
    assert np.array_equal(nlp.meta.x0, np.ones(41))
    nlp2 = M.GridapPDENLPModel(spec, device="cuda:0", memspace="host", lvar=-np.ones(41), uvar=np.ones(41))
    assert np.array_equal(nlp2.meta.lvar, -np.ones(41)) and np.array_equal(nlp2.meta.uvar, np.ones(41))
    with pytest.raises(M.DimensionError):
        M.GridapPDENLPModel(spec, device="cuda:0", memspace="host", lvar=np.zeros(40))
    with pytest.raises(M.DimensionError):
        M.GridapPDENLPModel(spec, device="cuda:0", memspace="host", x0=np.zeros(40))
    spec_k = pkg.burger1d_param(5)
    nlp3 = M.GridapPDENLPModel(spec_k, device="cuda:0", memspace="host", lvark=[0.0], uvark=[2.0])
    assert nlp3.meta.lvar[0] == 0.0 and nlp3.meta.uvar[0] == 2.0 and np.isinf(nlp3.meta.lvar[1:]).all()
