"""ctypes wrapper of the CPU oracle (oracle/pdenlp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never from the product
package.  Parity status: see the header of pdenlp_oracle.cpp ("parity unpinned"
for jac/hess values; numbering and KATs pinned by the reference's own tests).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_pdenlp.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pdenlp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_pdenlp.so"])
    return _LIB_PATH


class _Problem(C.Structure):
    _fields_ = [
        ("integrand", C.c_int32), ("dim", C.c_int32), ("order_y", C.c_int32), ("order_u", C.c_int32),
        ("degree", C.c_int32), ("nparam", C.c_int32), ("n", C.c_int32 * 3), ("nthreads", C.c_int32),
        ("lo", C.c_double * 3), ("hi", C.c_double * 3),
        ("ncells", C.c_int64), ("cell0", C.c_int64), ("nfree_y", C.c_int64), ("nfree_u", C.c_int64),
        ("cell_dofs_y", C.c_void_p), ("cell_dofs_u", C.c_void_p), ("dir_y", C.c_void_p), ("dir_u", C.c_void_p),
        ("params", C.c_double * 8),
        ("nfy", C.c_int32), ("nfu", C.c_int32), ("field_nfree", C.c_int64 * 8), ("field_ndir", C.c_int64 * 8),
        ("compat_shifted_x", C.c_int32), ("reserved", C.c_int32),
        ("vertex_coords", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Oracle:
    """CPU restatement of the reference callbacks for one ProblemSpec slab.

    ``spec`` is a pdenlp_b200.problems.ProblemSpec; only its mesh, dof maps,
    Dirichlet values, integrand id and scalar parameters are read — shape
    tables, quadrature and coefficient fields are restated inside the C++.
    """

    def __init__(self, spec, cell_range=None, nthreads: int = 0, compat_shifted_x: bool = False):
        m = spec.model
        c0, c1 = (0, m.ncells) if cell_range is None else cell_range
        self.spec = spec
        self._keep = []
        P = _Problem()
        P.integrand = spec.integrand
        P.compat_shifted_x = int(bool(compat_shifted_x))
        P.dim = m.dim
        if getattr(spec, "is_multifield", False):
            self._init_multifield(P, spec, c0, c1, nthreads)
            return
        P.order_y = spec.Ypde.order
        P.order_u = spec.Ycon.order if spec.Ycon is not None else 0
        P.degree = spec.degree
        P.nparam = spec.nparam
        for d in range(3):
            P.n[d] = m.n[d] if d < m.dim else 1
            P.lo[d] = m.lo[d] if d < m.dim else 0.0
            P.hi[d] = m.hi[d] if d < m.dim else 1.0
        P.nthreads = nthreads
        P.ncells = c1 - c0
        P.cell0 = c0
        P.nfree_y = spec.Ypde.nfree
        P.nfree_u = spec.Ycon.nfree if spec.Ycon is not None else 0
        cy = np.ascontiguousarray(spec.Ypde.cell_dof_ids[c0:c1], dtype=np.int32)
        dy = np.ascontiguousarray(spec.Ypde.dirichlet_vals, dtype=np.float64)
        self._keep += [cy, dy]
        P.cell_dofs_y = _p(cy)
        P.dir_y = _p(dy)
        if spec.Ycon is not None:
            cu = np.ascontiguousarray(spec.Ycon.cell_dof_ids[c0:c1], dtype=np.int32)
            du = np.ascontiguousarray(spec.Ycon.dirichlet_vals, dtype=np.float64)
            self._keep += [cu, du]
            P.cell_dofs_u = _p(cu)
            P.dir_u = _p(du)
        for i, v in enumerate(spec.params):
            P.params[i] = v
        self._finish_init(P, spec)

    def _init_multifield(self, P, spec, c0, c1, nthreads):
        m = spec.model
        Ys, Us = list(spec.Yfields), list(spec.Ufields)
        P.order_y = Ys[0].order
        P.order_u = Us[0].order if Us else 0
        P.degree = spec.degree
        P.nparam = 0
        for d in range(3):
            P.n[d] = m.n[d] if d < m.dim else 1
            P.lo[d] = m.lo[d] if d < m.dim else 0.0
            P.hi[d] = m.hi[d] if d < m.dim else 1.0
        P.nthreads = nthreads
        P.ncells = c1 - c0
        P.cell0 = c0
        P.nfree_y = sum(f.nfree for f in Ys)
        P.nfree_u = sum(f.nfree for f in Us)
        P.nfy, P.nfu = len(Ys), len(Us)
        for i, f in enumerate(Ys + Us):
            P.field_nfree[i] = f.nfree
            P.field_ndir[i] = f.ndirichlet
        cy = np.ascontiguousarray(np.concatenate([f.cell_dof_ids[c0:c1] for f in Ys], axis=1), dtype=np.int32)
        dy = np.ascontiguousarray(np.concatenate([f.dirichlet_vals for f in Ys]), dtype=np.float64)
        self._keep += [cy, dy]
        P.cell_dofs_y, P.dir_y = _p(cy), _p(dy)
        if Us:
            cu = np.ascontiguousarray(np.concatenate([f.cell_dof_ids[c0:c1] for f in Us], axis=1), dtype=np.int32)
            du = np.ascontiguousarray(np.concatenate([f.dirichlet_vals for f in Us]), dtype=np.float64)
            self._keep += [cu, du]
            P.cell_dofs_u, P.dir_u = _p(cu), _p(du)
        for i, v in enumerate(spec.params):
            P.params[i] = v
        self._finish_init(P, spec)

    def _finish_init(self, P, spec):
        if getattr(spec.model, "is_mapped", False):  # CartesianDiscreteModel(...; map = f): mapped vertex coordinates
            vc = np.ascontiguousarray(spec.model.vertex_coords(), dtype=np.float64)
            self._keep.append(vc)
            P.vertex_coords = _p(vc)
        self.P = P
        self.nvar = spec.nparam + P.nfree_y + P.nfree_u
        self.ncon = P.nfree_y
        sz = np.zeros(7, dtype=np.int64)
        self._check(lib().oracle_sizes(C.byref(P), _p(sz)))
        (self.nnzj_k, self.nnzj_y, self.nnzj_u, self.nnzh_k_obj, self.nnzh_fe, self.nnzh_k_con,
         self.nnzh_fe_obj) = (int(v) for v in sz)  # nnzh_fe_obj = 0 for NoFETerm objectives, else nnzh_fe
        self.nnzj = self.nnzj_k + self.nnzj_y + self.nnzj_u
        self.nnzh_obj = self.nnzh_k_obj + self.nnzh_fe_obj
        self.nnzh = self.nnzh_obj + self.nnzh_k_con + self.nnzh_fe

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise RuntimeError(f"oracle: unsupported configuration (rc={rc})")

    def set_threads(self, n: int):
        self.P.nthreads = n

    def obj(self, x):
        f = C.c_double()
        self._check(lib().oracle_obj(C.byref(self.P), _p(np.ascontiguousarray(x)), C.byref(f)))
        return f.value

    def grad(self, x):
        g = np.empty(self.nvar)
        self._check(lib().oracle_grad(C.byref(self.P), _p(np.ascontiguousarray(x)), _p(g)))
        return g

    def cons(self, x):
        c = np.empty(self.ncon)
        self._check(lib().oracle_cons(C.byref(self.P), _p(np.ascontiguousarray(x)), _p(c)))
        return c

    def jac_structure(self):
        r = np.empty(self.nnzj, dtype=np.int64); c = np.empty(self.nnzj, dtype=np.int64)
        self._check(lib().oracle_jac_structure(C.byref(self.P), _p(r), _p(c)))
        return r, c

    def jac_coord(self, x, out=None):
        v = np.empty(self.nnzj) if out is None else out
        self._check(lib().oracle_jac_coord(C.byref(self.P), _p(np.ascontiguousarray(x)), _p(v)))
        return v

    def hess_structure(self):
        r = np.empty(self.nnzh, dtype=np.int64); c = np.empty(self.nnzh, dtype=np.int64)
        self._check(lib().oracle_hess_structure(C.byref(self.P), _p(r), _p(c)))
        return r, c

    def hess_coord(self, x, lam=None, obj_weight=1.0, out=None):
        v = np.empty(self.nnzh) if out is None else out
        lp = _p(np.ascontiguousarray(lam)) if lam is not None else None
        self._check(lib().oracle_hess_coord(C.byref(self.P), _p(np.ascontiguousarray(x)), lp,
                                            C.c_double(obj_weight), _p(v)))
        return v

    # products exactly as the reference forms them: coo_prod!/coo_sym_prod! on the coord values
    def jprod(self, x, v):
        r, c = self.jac_structure(); vals = self.jac_coord(x)
        out = np.empty(self.ncon)
        lib().oracle_coo_prod(C.c_int64(vals.size), _p(r), _p(c), _p(vals), _p(np.ascontiguousarray(v)), C.c_int64(self.ncon), _p(out))
        return out

    def jtprod(self, x, v):
        r, c = self.jac_structure(); vals = self.jac_coord(x)
        out = np.empty(self.nvar)
        lib().oracle_coo_prod(C.c_int64(vals.size), _p(c), _p(r), _p(vals), _p(np.ascontiguousarray(v)), C.c_int64(self.nvar), _p(out))
        return out

    def hprod(self, x, v, lam=None, obj_weight=1.0):
        if lam is None and obj_weight == 0:
            return np.zeros(self.nvar)  # GridapPDENLPModel.jl:313-316
        r, c = self.hess_structure(); vals = self.hess_coord(x, lam, obj_weight)
        out = np.empty(self.nvar)
        lib().oracle_coo_sym_prod(C.c_int64(vals.size), _p(c), _p(r), _p(vals), _p(np.ascontiguousarray(v)), C.c_int64(self.nvar), _p(out))
        return out

    def tables(self):
        if getattr(self.spec, "is_multifield", False):
            ny = self.spec.Yfields[0].nloc; nu = self.spec.Ufields[0].nloc if self.spec.Ufields else 0
        else:
            ny = self.spec.Ypde.nloc; nu = self.spec.Ycon.nloc if self.spec.Ycon is not None else 0
        dim = self.spec.model.dim
        nqmax = 27
        nq = C.c_int32()
        xi = np.zeros(nqmax * dim); w = np.zeros(nqmax); phiy = np.zeros(nqmax * ny)
        dphiy = np.zeros(nqmax * ny * dim); phiu = np.zeros(nqmax * max(nu, 1))
        lib().oracle_tables(C.byref(self.P), C.byref(nq), _p(xi), _p(w), _p(phiy), _p(dphiy), _p(phiu))
        q = nq.value
        return (xi[: q * dim].reshape(q, dim), w[:q], phiy[: q * ny].reshape(q, ny),
                dphiy[: q * ny * dim].reshape(q, ny, dim), phiu[: q * nu].reshape(q, nu))


def max_threads() -> int:
    return int(lib().oracle_max_threads())
