"""Host-side mirror of the reference interface for the hot path.

``GridapPDENLPModel`` here exposes the NLPModels callbacks of the reference's
``GridapPDENLPModel`` (/root/reference/src/GridapPDENLPModel.jl:219-661) with the
same names, argument meaning, error behaviour (``DimensionError`` before any
work — the reference's ``@lencheck``) and ``Counters`` bookkeeping
(increment own counter; nested evaluations do not count:
src/GridapPDENLPModel.jl:319-320,353-354,426-427,483-484,515-516).  Julia's
``f!`` becomes ``f_`` (in-place, torch convention); ``f`` allocates.

Every callback body is exactly one call into libpdenlp_cuda.so.  Vectors are
torch CUDA float64 tensors (memspace "device") or numpy / torch CPU arrays
(memspace "host": the library stages them, copies inside the call).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import capi
from .flatten import FlatModel, flatten
from .flatten import hess_structure as _host_hess_structure  # noqa: F401 (re-exported for tests)
from .flatten import jac_structure as _host_jac_structure  # noqa: F401
from .problems import ProblemSpec


from .model_errors import DimensionError  # noqa: E402,F401  (NLPModels.DimensionError)


_COUNTERS = (
    "neval_obj", "neval_grad", "neval_cons", "neval_cons_lin", "neval_cons_nln", "neval_jcon", "neval_jgrad",
    "neval_jac", "neval_jac_lin", "neval_jac_nln", "neval_jprod", "neval_jprod_lin", "neval_jprod_nln",
    "neval_jtprod", "neval_jtprod_lin", "neval_jtprod_nln", "neval_hess", "neval_hprod", "neval_jhess", "neval_jhprod",
)


class Counters:
    def __init__(self):
        self.reset()

    def reset(self):
        for c in _COUNTERS:
            setattr(self, c, 0)

    def as_dict(self):
        return {c: getattr(self, c) for c in _COUNTERS}


@dataclass
class NLPModelMeta:
    nvar: int
    ncon: int
    nnzj: int
    nnzh: int
    lin_nnzj: int = 0
    nln_nnzj: int = 0
    nlin: int = 0
    nnln: int = 0
    name: str = "Generic"
    minimize: bool = True
    x0: Optional[np.ndarray] = None
    lvar: Optional[np.ndarray] = None
    uvar: Optional[np.ndarray] = None
    lcon: Optional[np.ndarray] = None
    ucon: Optional[np.ndarray] = None
    y0: Optional[np.ndarray] = None


@dataclass
class PDENLPMeta:
    """The fields of the reference's PDENLPMeta that exist without Gridap
    (src/GridapPDENLPModel.jl:24-49)."""
    nvar_pde: int
    nvar_con: int
    nparam: int
    nnzh_obj: int
    nnzj_k: int
    nnzj_y: int
    nnzj_u: int
    nnzh_fe: int
    nnzh_k_obj: int
    nnzh_k_con: int


def _is_torch(a):
    return isinstance(a, torch.Tensor)


class GridapPDENLPModel:
    """Drop-in mirror of the reference model for one cell slab on one GPU.

    Parameters
    ----------
    spec : ProblemSpec — what the Gridap user code describes (mesh, spaces, f, res).
    device : torch device string.
    cell_range : (c0, c1) slab of cells owned by this handle (default: all).
    memspace : "device" (vectors are CUDA tensors, used in place) or "host".
    deterministic : sum the dof-indexed outputs in a fixed order (coloured launches, no atomics): bitwise reproducible.
    templates : False evaluates every cell with the generic kernels (PDENLP_FLAG_NO_TEMPLATES) instead of copying the
        image of a template cell (affine problems on identical cells): same values, used to compare both paths.
    """

    def __init__(self, spec: ProblemSpec, device: str = "cuda:0", cell_range=None, memspace: str = "device",
                 with_coef: bool = True, flat: Optional[FlatModel] = None, geometry: Optional[str] = None,
                 deterministic: bool = False, templates: bool = True, x0=None, lvar=None, uvar=None,
                 lvary=None, uvary=None, lvaru=None, uvaru=None, lvark=None, uvark=None):
        self.spec = spec
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("GridapPDENLPModel needs a CUDA device: the hot path has no CPU implementation")
        self.memspace = memspace
        fm = flat if flat is not None else flatten(spec, cell_range, with_coef=with_coef, geometry=geometry)
        self.flat = fm
        ms = capi.MEM_DEVICE if memspace == "device" else capi.MEM_HOST
        self.handle = capi.Handle(fm, device=self.device.index if self.device.index is not None else 0,
                                  memspace=ms, with_coef=with_coef,
                                  flags=(capi.FLAG_DETERMINISTIC if deterministic else 0) | (0 if templates else capi.FLAG_NO_TEMPLATES))
        I = self.handle.info
        self.affine = bool(getattr(spec, "affine", False))
        self._jac_plan = self._hess_plan = None
        self._Alin_vals = None
        if self.affine:
            # AffineFEOperator: lin = 1:ncon, lin_nnzj = nnzj = nnz(A), nln_nnzj = 0
            # (src/additional_constructors.jl:213-223); Jacobian structure = findnz(A)
            # (src/jacobian_functions.jl:11-21); empty constraint-Hessian structure
            # (src/hessian_struct_nnzh_functions.jl:70-72) => nnzh = nnzh_obj.
            if spec.nparam:
                raise NotImplementedError("AffineFEOperator with algebraic variables (the reference only warns)")
            if memspace != "device":
                raise NotImplementedError("the linear-API model keeps its vectors on the device")
            self._jac_plan = self._make_plan(False, I.nnzj)
            nnzj, nnzh = self._jac_plan.nnz, I.nnzh_obj
            lin_nnzj, nln_nnzj, nlin, nnln = nnzj, 0, I.ncon, 0
        else:
            nnzj, nnzh = I.nnzj, I.nnzh
            lin_nnzj, nln_nnzj, nlin, nnln = 0, I.nnzj, 0, I.ncon
        self.unconstrained = bool(getattr(spec, "unconstrained", False))
        ncon = I.ncon
        if self.unconstrained:
            # the reference's unconstrained constructor (src/additional_constructors.jl:1-66): no operator,
            # ncon = 0, empty Jacobian, Hessian = objective segment only
            ncon, nnzj, nnzh, lin_nnzj, nln_nnzj, nlin, nnln = 0, 0, I.nnzh_obj, 0, 0, 0, 0
        lvar, uvar = self._bounds(spec, fm, I.nvar, lvar, uvar, lvary, uvary, lvaru, uvaru, lvark, uvark)
        x0 = np.zeros(I.nvar) if x0 is None else np.asarray(x0, dtype=np.float64)
        if x0.shape != (I.nvar,):
            raise DimensionError("x0", I.nvar, x0.shape[0] if x0.ndim == 1 else tuple(x0.shape))
        self.meta = NLPModelMeta(
            nvar=I.nvar, ncon=ncon, nnzj=nnzj, nnzh=nnzh, lin_nnzj=lin_nnzj, nln_nnzj=nln_nnzj, nlin=nlin, nnln=nnln,
            name=spec.name, x0=x0, lvar=lvar, uvar=uvar,
            lcon=np.zeros(ncon), ucon=np.zeros(ncon), y0=np.zeros(ncon),
        )
        self.pdemeta = PDENLPMeta(
            nvar_pde=fm.nfree_y, nvar_con=fm.nfree_u, nparam=fm.nparam, nnzh_obj=I.nnzh_obj, nnzj_k=I.nnzj_k,
            nnzj_y=I.nnzj_y, nnzj_u=I.nnzj_u, nnzh_fe=I.nnzh_fe, nnzh_k_obj=I.nnzh_k_obj, nnzh_k_con=I.nnzh_k_con,
        )
        self.counters = Counters()
        self._use_current_stream()

    @staticmethod
    def _bounds(spec, fm, nvar, lvar, uvar, lvary, uvary, lvaru, uvaru, lvark, uvark):
        """lvar / uvar as the constrained constructor builds them (src/additional_constructors.jl:178-204):
        whole vectors, or [kappa bounds ; bounds_functions_to_vectors(y and u bounds as functions or vectors)];
        the default is unbounded."""
        p = fm.nparam
        if lvar is not None or uvar is not None:
            lvar = np.full(nvar, -np.inf) if lvar is None else np.asarray(lvar, dtype=np.float64)
            uvar = np.full(nvar, np.inf) if uvar is None else np.asarray(uvar, dtype=np.float64)
            for name, v in (("lvar", lvar), ("uvar", uvar)):
                if v.shape != (nvar,):
                    raise DimensionError(name, nvar, v.shape[0] if v.ndim == 1 else tuple(v.shape))
            return lvar, uvar
        if all(b is None for b in (lvary, uvary, lvaru, uvaru, lvark, uvark)):
            return np.full(nvar, -np.inf), np.full(nvar, np.inf)
        from .bounds import bounds_functions_to_vectors
        if getattr(spec, "is_multifield", False):
            Ypde, Ycon = list(spec.Yfields), (list(spec.Ufields) or None)
        else:
            Ypde, Ycon = spec.Ypde, spec.Ycon
        ny, nu = fm.nfree_y, fm.nfree_u
        lvary = np.full(ny, -np.inf) if lvary is None else lvary
        uvary = np.full(ny, np.inf) if uvary is None else uvary
        lvaru = np.full(nu, -np.inf) if lvaru is None else lvaru
        uvaru = np.full(nu, np.inf) if uvaru is None else uvaru
        lyu, uyu = bounds_functions_to_vectors(Ycon, Ypde, spec.model, lvary, uvary, lvaru, uvaru)
        lk = np.full(p, -np.inf) if lvark is None else np.asarray(lvark, dtype=np.float64)
        uk = np.full(p, np.inf) if uvark is None else np.asarray(uvark, dtype=np.float64)
        for name, v in (("lvark", lk), ("uvark", uk)):
            if v.shape != (p,):
                raise DimensionError(name, p, v.shape[0] if v.ndim == 1 else tuple(v.shape))
        return np.concatenate([lk, lyu]), np.concatenate([uk, uyu])

    # ------------------------------------------------------------------ plumbing
    def _use_current_stream(self):
        self.handle.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        self.handle.close()

    def launch_count(self) -> int:
        return self.handle.launch_count()

    def staged_input_bytes(self, vectors=("x",)) -> int:
        """Bytes a host-memspace model copies to the device for the given input vectors ("x" / "v": x-layout,
        "lambda": constraint layout): only the dof ranges its cells touch are staged."""
        ylo, yhi, ulo, uhi = self.handle.input_ranges()
        p = int(self.spec.nparam)
        per = {"x": p + (yhi - ylo) + (uhi - ulo), "v": p + (yhi - ylo) + (uhi - ulo), "lambda": yhi - ylo}
        return 8 * sum(per[v] for v in vectors)

    def _new(self, n, like=None, dtype=torch.float64):
        if self.memspace == "device":
            return torch.empty(n, dtype=dtype, device=self.device)
        return np.empty(n, dtype=np.float64 if dtype == torch.float64 else np.int64)

    def _lencheck(self, n, **vecs):
        for name, v in vecs.items():
            if v.shape[0] != n or v.ndim != 1:
                raise DimensionError(name, n, v.shape[0] if v.ndim == 1 else tuple(v.shape))

    def _in(self, v):
        """Pointer of an input vector (contiguous float64 in the model's memory space)."""
        if v is None:
            return None, None
        if self.memspace == "device":
            if not _is_torch(v) or v.device != self.device:
                v = torch.as_tensor(np.asarray(v.cpu() if _is_torch(v) else v), dtype=torch.float64).to(self.device)
            if v.dtype != torch.float64 or not v.is_contiguous():
                v = v.to(torch.float64).contiguous()  # views with stride != 1 are staged
            return v.data_ptr(), v
        a = v.detach().cpu().numpy() if _is_torch(v) else np.asarray(v)
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a.ctypes.data, a

    def _out(self, v):
        """(pointer, buffer, needs_copy_back) for an output vector."""
        if self.memspace == "device":
            if not _is_torch(v) or v.device != self.device or v.dtype != torch.float64:
                raise TypeError("output must be a float64 CUDA tensor on the model's device")
            if v.is_contiguous():
                return v.data_ptr(), v, False
            tmp = torch.empty(v.shape, dtype=v.dtype, device=v.device)
            return tmp.data_ptr(), tmp, True
        if _is_torch(v):
            if v.is_contiguous() and v.dtype == torch.float64:
                return v.data_ptr(), v, False
            tmp = torch.empty(v.shape, dtype=torch.float64)
            return tmp.data_ptr(), tmp, True
        if v.flags.c_contiguous and v.dtype == np.float64:
            return v.ctypes.data, v, False
        tmp = np.empty(v.shape, dtype=np.float64)
        return tmp.ctypes.data, tmp, True

    @staticmethod
    def _finish(user, buf, copy_back):
        if copy_back:
            if _is_torch(user):
                user.copy_(buf if _is_torch(buf) else torch.from_numpy(buf))
            else:
                user[...] = buf
        return user

    # ------------------------------------------------------------------ compaction plans (COO -> CSC)
    def _make_plan(self, hess: bool, nnz: int):
        """Device plan summing duplicates of the COO structure into CSC order (findnz order)."""
        if hess:  # the model-level structure (objective segment only for unconstrained models)
            rows, cols = self.hess_structure()
            rows, cols = rows.contiguous(), cols.contiguous()
        else:
            rows = torch.empty(nnz, dtype=torch.int64, device=self.device)
            cols = torch.empty(nnz, dtype=torch.int64, device=self.device)
            self.handle.jac_structure(rows.data_ptr(), cols.data_ptr())
        I = self.handle.info
        with torch.cuda.device(self.device):
            plan = capi.Compaction(rows, cols, I.nvar if hess else I.ncon, I.nvar, capi.ORDER_CSC)
        return plan

    def _plan_structure(self, plan, with_ptr: bool = False):
        rows = torch.empty(plan.nnz, dtype=torch.int64, device=self.device)
        cols = torch.empty(plan.nnz, dtype=torch.int64, device=self.device)
        ptr = torch.empty(plan.ncols + 1, dtype=torch.int64, device=self.device) if with_ptr else None
        plan.structure(rows.data_ptr(), cols.data_ptr(), ptr.data_ptr() if with_ptr else None, on_device=True)
        return (rows, cols, ptr) if with_ptr else (rows, cols)

    def jac_csc(self, x):
        """NLPModels ``jac(nlp, x)`` as (colptr, rowval, nzval), 1-based like SparseMatrixCSC:
        sparse(rows, cols, vals, ncon, nvar) with duplicates summed on the device."""
        if self.memspace != "device":
            raise NotImplementedError("jac_csc needs a device-memspace model")
        I = self.handle.info
        if self._jac_plan is None:
            self._jac_plan = self._make_plan(False, I.nnzj)
        coo = torch.empty(I.nnzj, dtype=torch.float64, device=self.device)
        self._use_current_stream()
        px, kx = self._in(x)
        self.handle.jac_coord(px, coo.data_ptr())
        out = torch.empty(self._jac_plan.nnz, dtype=torch.float64, device=self.device)
        self._jac_plan.apply(coo.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        rows, cols, ptr = self._plan_structure(self._jac_plan, with_ptr=True)
        return ptr, rows, out

    def hess_csc(self, x, y=None, obj_weight: float = 1.0):
        """NLPModels ``hess(nlp, x[, y])``: lower triangle sparse(rows, cols, vals, nvar, nvar) with
        the overlapping objective / constraint segments summed, as (colptr, rowval, nzval)."""
        if self.memspace != "device" or self.affine:
            raise NotImplementedError
        if self._hess_plan is None:
            self._hess_plan = self._make_plan(True, self.meta.nnzh)
        coo = self.hess_coord(x, y, obj_weight=obj_weight)
        out = torch.empty(self._hess_plan.nnz, dtype=torch.float64, device=self.device)
        self._hess_plan.apply(coo.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        rows, cols, ptr = self._plan_structure(self._hess_plan, with_ptr=True)
        return ptr, rows, out

    # ------------------------------------------------------------------ linear API (AffineFEOperator)
    def _require_affine(self):
        if not self.affine:
            raise TypeError("linear-API methods exist only for AffineFEOperator models")

    def cons_lin_(self, x, c):
        """cons_lin!: mul!(c, A, x); c -= b (src/GridapPDENLPModel.jl:460-471) — evaluated matrix-free
        as the assembled residual of the same bilinear/linear forms."""
        self._require_affine()
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.nlin, c=c)
        self.counters.neval_cons_lin += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pc, bc, cb = self._out(c)
        self.handle.cons(px, pc)
        return self._finish(c, bc, cb)

    def _lin_vals(self):
        if self._Alin_vals is None:  # A is constant: assemble its CSC values once
            I = self.handle.info
            coo = torch.empty(I.nnzj, dtype=torch.float64, device=self.device)
            x0 = torch.zeros(I.nvar, dtype=torch.float64, device=self.device)
            self.handle.jac_coord(x0.data_ptr(), coo.data_ptr())
            out = torch.empty(self._jac_plan.nnz, dtype=torch.float64, device=self.device)
            self._jac_plan.apply(coo.data_ptr(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
            self._Alin_vals = out
        return self._Alin_vals

    def jac_lin_coord_(self, x, vals):
        """jac_lin_coord!: vals .= findnz(get_matrix(op))[3] (src/GridapPDENLPModel.jl:597-608)."""
        self._require_affine()
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.lin_nnzj, vals=vals)
        self.counters.neval_jac_lin += 1
        self._use_current_stream()
        vals.copy_(self._lin_vals())
        return vals

    def jac_lin_structure_(self, rows, cols):
        self._require_affine()
        self._lencheck(self.meta.lin_nnzj, rows=rows, cols=cols)
        r, c = self._plan_structure(self._jac_plan)
        rows.copy_(r); cols.copy_(c)
        return rows, cols

    def jprod_lin_(self, x, v, Jv):
        self._require_affine()
        self._lencheck(self.meta.nvar, x=x, v=v)
        self._lencheck(self.meta.nlin, Jv=Jv)
        self.counters.neval_jprod_lin += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pv, kv = self._in(v)
        po, bo, cb = self._out(Jv)
        self.handle.jprod(px, pv, po)
        return self._finish(Jv, bo, cb)

    def jtprod_lin_(self, x, v, Jtv):
        self._require_affine()
        self._lencheck(self.meta.nvar, x=x, Jtv=Jtv)
        self._lencheck(self.meta.nlin, v=v)
        self.counters.neval_jtprod_lin += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pv, kv = self._in(v)
        po, bo, cb = self._out(Jtv)
        self.handle.jtprod(px, pv, po)
        return self._finish(Jtv, bo, cb)

    # ------------------------------------------------------------------ obj / grad / cons
    def obj(self, x) -> float:
        self._lencheck(self.meta.nvar, x=x)
        self.counters.neval_obj += 1
        self._use_current_stream()
        px, kx = self._in(x)
        return self.handle.obj(px)

    def grad_(self, x, g):
        self._lencheck(self.meta.nvar, x=x, g=g)
        self.counters.neval_grad += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pg, bg, cb = self._out(g)
        self.handle.grad(px, pg)
        return self._finish(g, bg, cb)

    def grad(self, x):
        return self.grad_(x, self._new(self.meta.nvar))

    def cons_nln_(self, x, c):
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.nnln, c=c)
        self.counters.neval_cons_nln += 1
        if self.unconstrained:
            return c
        self._use_current_stream()
        px, kx = self._in(x)
        pc, bc, cb = self._out(c)
        self.handle.cons(px, pc)
        return self._finish(c, bc, cb)

    def cons_(self, x, c):
        """NLPModels generic cons!: nlin = 0 for FEOperatorFromWeakForm models
        (src/additional_constructors.jl:213-223), so this is cons_nln! on the whole vector."""
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.ncon, c=c)
        self.counters.neval_cons += 1
        return self.cons_lin_(x, c) if self.meta.nlin > 0 else self.cons_nln_(x, c)

    def cons(self, x):
        return self.cons_(x, self._new(self.meta.ncon))

    # ------------------------------------------------------------------ Jacobian
    def jac_nln_coord_(self, x, vals):
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.nln_nnzj, vals=vals)
        self.counters.neval_jac_nln += 1
        if self.unconstrained:
            return vals
        self._use_current_stream()
        px, kx = self._in(x)
        pv, bv, cb = self._out(vals)
        self.handle.jac_coord(px, pv)
        return self._finish(vals, bv, cb)

    def jac_coord_(self, x, vals):
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.nnzj, vals=vals)
        self.counters.neval_jac += 1
        return self.jac_lin_coord_(x, vals) if self.meta.nlin > 0 else self.jac_nln_coord_(x, vals)

    def jac_coord(self, x):
        return self.jac_coord_(x, self._new(self.meta.nnzj))

    def jac_nln_structure_(self, rows, cols):
        self._lencheck(self.meta.nln_nnzj, rows=rows, cols=cols)
        if self.unconstrained:
            return rows, cols
        return self._structure(False, rows, cols)

    def jac_structure_(self, rows, cols):
        return self.jac_lin_structure_(rows, cols) if self.meta.nlin > 0 else self.jac_nln_structure_(rows, cols)

    def jac_structure(self):
        return self.jac_structure_(self._new(self.meta.nnzj, dtype=torch.int64), self._new(self.meta.nnzj, dtype=torch.int64))

    def _structure(self, hess, rows, cols):
        self._use_current_stream()
        ptrs = []
        for v in (rows, cols):
            if _is_torch(v):
                if v.dtype != torch.int64 or not v.is_contiguous():
                    raise TypeError("structure outputs must be contiguous int64")
                ptrs.append(v.data_ptr())
            else:
                if v.dtype != np.int64 or not v.flags.c_contiguous:
                    raise TypeError("structure outputs must be contiguous int64")
                ptrs.append(v.ctypes.data)
        (self.handle.hess_structure if hess else self.handle.jac_structure)(ptrs[0], ptrs[1])
        return rows, cols

    def jprod_nln_(self, x, v, Jv):
        self._lencheck(self.meta.nvar, x=x, v=v)
        self._lencheck(self.meta.nnln, Jv=Jv)
        self.counters.neval_jprod_nln += 1
        if self.unconstrained:
            return Jv
        self._use_current_stream()
        px, kx = self._in(x)
        pv, kv = self._in(v)
        po, bo, cb = self._out(Jv)
        self.handle.jprod(px, pv, po)
        return self._finish(Jv, bo, cb)

    def jprod_(self, x, v, Jv):
        self._lencheck(self.meta.nvar, x=x, v=v)
        self._lencheck(self.meta.ncon, Jv=Jv)
        self.counters.neval_jprod += 1
        return self.jprod_lin_(x, v, Jv) if self.meta.nlin > 0 else self.jprod_nln_(x, v, Jv)

    def jprod(self, x, v):
        return self.jprod_(x, v, self._new(self.meta.ncon))

    def jtprod_nln_(self, x, v, Jtv):
        self._lencheck(self.meta.nvar, x=x, Jtv=Jtv)
        self._lencheck(self.meta.nnln, v=v)
        self.counters.neval_jtprod_nln += 1
        if self.unconstrained:
            Jtv[...] = 0.0
            return Jtv
        self._use_current_stream()
        px, kx = self._in(x)
        pv, kv = self._in(v)
        po, bo, cb = self._out(Jtv)
        self.handle.jtprod(px, pv, po)
        return self._finish(Jtv, bo, cb)

    def jtprod_(self, x, v, Jtv):
        self._lencheck(self.meta.nvar, x=x, Jtv=Jtv)
        self._lencheck(self.meta.ncon, v=v)
        self.counters.neval_jtprod += 1
        return self.jtprod_lin_(x, v, Jtv) if self.meta.nlin > 0 else self.jtprod_nln_(x, v, Jtv)

    def jtprod(self, x, v):
        return self.jtprod_(x, v, self._new(self.meta.nvar))

    # ------------------------------------------------------------------ Hessian
    def hess_coord_(self, x, *args, obj_weight: float = 1.0):
        """hess_coord!(nlp, x, vals; obj_weight) or hess_coord!(nlp, x, y, vals; obj_weight)."""
        if len(args) == 1:
            lam, vals = None, args[0]
        elif len(args) == 2:
            lam, vals = args
        else:
            raise TypeError("hess_coord_(x, vals) or hess_coord_(x, y, vals)")
        self._lencheck(self.meta.nvar, x=x)
        if lam is not None:
            self._lencheck(self.meta.ncon, y=lam)
        self._lencheck(self.meta.nnzh, vals=vals)
        self.counters.neval_hess += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pl, kl = self._in(lam)
        pv, bv, cb = self._out(vals)
        if self.affine or self.unconstrained:  # empty constraint-Hessian structure
            self.handle.hess_coord_obj(px, obj_weight, pv)
        else:
            self.handle.hess_coord(px, pl, obj_weight, pv)
        return self._finish(vals, bv, cb)

    def hess_coord(self, x, y=None, obj_weight: float = 1.0):
        vals = self._new(self.meta.nnzh)
        return self.hess_coord_(x, vals, obj_weight=obj_weight) if y is None else self.hess_coord_(x, y, vals, obj_weight=obj_weight)

    def hess_structure_(self, rows, cols):
        self._lencheck(self.meta.nnzh, rows=rows, cols=cols)
        if self.affine or self.unconstrained:  # objective segment only
            n = self.handle.info.nnzh
            r = self._new(n, dtype=torch.int64)
            c = self._new(n, dtype=torch.int64)
            self._structure(True, r, c)
            if _is_torch(rows):
                rows.copy_(r[: self.meta.nnzh]); cols.copy_(c[: self.meta.nnzh])
            else:
                rows[:] = r[: self.meta.nnzh]; cols[:] = c[: self.meta.nnzh]
            return rows, cols
        return self._structure(True, rows, cols)

    def hess_structure(self):
        return self.hess_structure_(self._new(self.meta.nnzh, dtype=torch.int64), self._new(self.meta.nnzh, dtype=torch.int64))

    def hprod_(self, x, *args, obj_weight: float = 1.0):
        """hprod!(nlp, x, v, Hv; obj_weight) or hprod!(nlp, x, y, v, Hv; obj_weight)."""
        if len(args) == 2:
            lam, (v, Hv) = None, args
        elif len(args) == 3:
            lam, v, Hv = args
        else:
            raise TypeError("hprod_(x, v, Hv) or hprod_(x, y, v, Hv)")
        self._lencheck(self.meta.nvar, x=x, v=v, Hv=Hv)
        if lam is not None:
            self._lencheck(self.meta.ncon, y=lam)
        self.counters.neval_hprod += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pl, kl = self._in(lam)
        pv, kv = self._in(v)
        po, bo, cb = self._out(Hv)
        self.handle.hprod(px, None if (self.affine or self.unconstrained) else pl, obj_weight, pv, po)
        return self._finish(Hv, bo, cb)

    def hprod(self, x, *args, obj_weight: float = 1.0):
        Hv = self._new(self.meta.nvar)
        return self.hprod_(x, *args, Hv, obj_weight=obj_weight)

    # ------------------------------------------------------------------ NLPModels generic wrappers
    # (the reference inherits these from NLPModels; they compose the callbacks above)
    def objgrad_(self, x, g):
        """objgrad!(nlp, x, g) = (obj(nlp, x), grad!(nlp, x, g)); both from one sweep over the cells
        (pdenlp_objgrad), counted as one obj and one grad evaluation like the NLPModels default."""
        self._lencheck(self.meta.nvar, x=x, g=g)
        self.counters.neval_obj += 1
        self.counters.neval_grad += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pg, bg, cb = self._out(g)
        f = self.handle.objgrad(px, pg)
        return f, self._finish(g, bg, cb)

    def objgrad(self, x):
        return self.objgrad_(x, self._new(self.meta.nvar, like=x))

    def objcons_(self, x, c):
        """objcons!(nlp, x, c) = (obj(nlp, x), cons!(nlp, x, c)); for a plain FEOperatorFromWeakForm model both
        come from one sweep over the cells (pdenlp_objcons), counted like the NLPModels default."""
        if self.meta.ncon == 0 or self.meta.nlin > 0:
            return self.obj(x), (self.cons_(x, c) if self.meta.ncon > 0 else c)
        self._lencheck(self.meta.nvar, x=x)
        self._lencheck(self.meta.ncon, c=c)
        self.counters.neval_obj += 1
        self.counters.neval_cons += 1
        self.counters.neval_cons_nln += 1
        self._use_current_stream()
        px, kx = self._in(x)
        pc, bc, cb = self._out(c)
        f = self.handle.objcons(px, pc)
        return f, self._finish(c, bc, cb)

    def objcons(self, x):
        return self.objcons_(x, self._new(self.meta.ncon, like=x))

    def jac(self, x):
        """jac(nlp, x): the Jacobian as (colptr, rowval, nzval) of a SparseMatrixCSC (see jac_csc)."""
        return self.jac_csc(x)

    def hess(self, x, y=None, obj_weight: float = 1.0):
        """hess(nlp, x[, y]; obj_weight): lower triangle as (colptr, rowval, nzval) (see hess_csc)."""
        return self.hess_csc(x, y, obj_weight=obj_weight)

    # ------------------------------------------------------------------ operators (jac_op!/hess_op!)
    def jac_op(self, x):
        """Returns (prod, tprod) closures: v -> J(x) v and w -> J(x)^T w
        (src/GridapPDENLPModel.jl:610-634; matrix-free here)."""
        return (lambda v: self.jprod(x, v)), (lambda w: self.jtprod(x, w))

    def hess_op(self, x, y=None, obj_weight: float = 1.0):
        if y is None:
            return lambda v: self.hprod(x, v, obj_weight=obj_weight)
        return lambda v: self.hprod(x, y, v, obj_weight=obj_weight)

    def reset_(self):
        self.counters.reset()
        return self
