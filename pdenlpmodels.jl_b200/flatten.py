"""Flattening layer: FE objects -> the plain arrays of ``pdenlp_desc``.

This is the Python stand-in for the Julia layer described in the north star
("a Julia layer flattens these into device arrays: cell-to-dof maps,
reference-FE shape-function and gradient tables at quadrature points, weights
and Jacobians"); julia/PDENLPCuda.jl holds the Julia version that reads the
same data from Gridap objects.  It also restates, vectorised, the COO structure
the reference computes once in its constructor
(/root/reference/src/additional_constructors.jl:206-211,259-260) so that the
container has a structure source without Gridap.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .cartesian import gauss_legendre_01, lagrange_tables
from .problems import ProblemSpec, KAPPA_POISSON


@dataclass
class FlatModel:
    """Host arrays of one cell slab, laid out as pdenlp_desc expects."""
    spec: ProblemSpec
    cell0: int
    ncells: int
    dim: int
    ny: int
    nu: int
    nq: int
    nparam: int
    nfree_y: int
    nfree_u: int
    cell_dofs_y: np.ndarray  # (ncells, ny) int32
    cell_dofs_u: np.ndarray  # (ncells, nu) int32
    dir_y: np.ndarray
    dir_u: np.ndarray
    phi_y: np.ndarray   # (nq, ny)
    grad_y: np.ndarray  # (nq, ny, dim) physical gradients
    phi_u: np.ndarray   # (nq, nu)
    wdet: np.ndarray    # (nq,)
    coef: np.ndarray    # (ncoef, ncells, nq)
    params: np.ndarray  # (8,)
    # multi-field layout: nfy state fields x ny local dofs, nfu control fields x nu;
    # cell_dofs_y is (ncells, nfy*ny) field-major, ids are per-field (no multi-field offset);
    # field_nfree / field_ndir list the state fields first, then the control fields
    nfy: int = 1
    nfu: int = 1
    field_nfree: np.ndarray = None
    field_ndir: np.ndarray = None
    # per-cell geometry (pdenlp_desc ABI v3): None = uniform fast path (grad_y physical, wdet = w * detJ); otherwise
    # grad_y / wdet are the REFERENCE gradients / weights and cell_jinvT (ncells, geom_nq, dim, dim), cell_detj
    # (ncells, geom_nq) hold J^-T and |det J| of every cell map (geom_nq = 1: affine cells, = nq: per point)
    cell_jinvT: Optional[np.ndarray] = None
    cell_detj: Optional[np.ndarray] = None
    geom_nq: int = 0

    @property
    def nvar(self) -> int:
        return self.nparam + self.nfree_y + self.nfree_u

    @property
    def ncon(self) -> int:
        return self.nfree_y


def _geometry(m, cells, pts, wts, dphi_ref, geometry):
    """Tables and per-cell geometry of a slab.  geometry: None (automatic: per-cell arrays only for a mapped mesh),
    "uniform" (the fast path; error for a mapped mesh), "per_point" (J^-T and det J at every quadrature point of every
    cell, the general case) or "affine" (one per cell, evaluated at the cell centre: exact for parallelotope cells).
    Returns (grad_y, wdet, cell_jinvT, cell_detj, geom_nq)."""
    if geometry is None:
        geometry = "per_point" if m.is_mapped else "uniform"
    if geometry == "uniform":
        if m.is_mapped:
            raise ValueError("a mapped mesh has no uniform geometry")
        return dphi_ref / m.h, wts * float(np.prod(m.h)), None, None, 0
    if geometry == "per_point":
        J, det = m.cell_jacobians(cells, pts)
    elif geometry == "affine":
        J, det = m.cell_jacobians(cells, np.full((1, m.dim), 0.5))
    else:
        raise ValueError("geometry must be None, 'uniform', 'per_point' or 'affine'")
    jinvT = np.ascontiguousarray(np.swapaxes(np.linalg.inv(J), -1, -2))  # (J^-1)^T: grad_x = jinvT @ grad_xi
    return dphi_ref, wts, jinvT, np.ascontiguousarray(det), J.shape[1]


def _flatten_mf(spec, cell_range, with_coef, geometry=None):
    m = spec.model
    c0, c1 = (0, m.ncells) if cell_range is None else cell_range
    Ys, Us = list(spec.Yfields), list(spec.Ufields)
    assert len({(f.order, f.nloc) for f in Ys}) == 1 and len({(f.order, f.nloc) for f in Us}) <= 1
    pts, wts = gauss_legendre_01(spec.degree, m.dim)
    nq = pts.shape[0]
    phi_y, dphi_y = lagrange_tables(Ys[0].order, m.dim, pts)
    phi_u = lagrange_tables(Us[0].order, m.dim, pts)[0] if Us else np.zeros((nq, 0))
    ncells = c1 - c0
    cells = np.arange(c0, c1, dtype=np.int64)
    grad_y, wdet, jinvT, detj, geom_nq = _geometry(m, cells, pts, wts, dphi_y, geometry)
    if with_coef:
        XQ = m.physical_points(cells, pts)
        coef = np.zeros((2, ncells, nq))
        for q in range(nq):
            X = XQ[:, q, :]
            if spec.target is not None:
                coef[0, :, q] = spec.target(X)
            if spec.source is not None:
                coef[1, :, q] = spec.source(X)
    else:
        coef = np.zeros((2, 0, nq))
    params = np.zeros(8)
    params[: len(spec.params)] = spec.params
    cat = lambda fs: (np.ascontiguousarray(np.concatenate([f.cell_dof_ids[c0:c1] for f in fs], axis=1))
                      if fs else np.zeros((ncells, 0), np.int32))
    catd = lambda fs: (np.ascontiguousarray(np.concatenate([f.dirichlet_vals for f in fs]).astype(np.float64))
                       if fs else np.zeros(0))
    return FlatModel(
        spec=spec, cell0=c0, ncells=ncells, dim=m.dim, ny=Ys[0].nloc, nu=(Us[0].nloc if Us else 0), nq=nq, nparam=0,
        nfree_y=sum(f.nfree for f in Ys), nfree_u=sum(f.nfree for f in Us),
        cell_dofs_y=cat(Ys), cell_dofs_u=cat(Us), dir_y=catd(Ys), dir_u=catd(Us),
        phi_y=np.ascontiguousarray(phi_y), grad_y=np.ascontiguousarray(grad_y), phi_u=np.ascontiguousarray(phi_u),
        wdet=np.ascontiguousarray(wdet), coef=np.ascontiguousarray(coef), params=params,
        nfy=len(Ys), nfu=len(Us),
        field_nfree=np.array([f.nfree for f in Ys + Us], dtype=np.int64),
        field_ndir=np.array([f.ndirichlet for f in Ys + Us], dtype=np.int64),
        cell_jinvT=jinvT, cell_detj=detj, geom_nq=geom_nq,
    )


def flatten(spec: ProblemSpec, cell_range: Optional[Tuple[int, int]] = None, with_coef: bool = True,
            geometry: Optional[str] = None) -> FlatModel:
    """``geometry``: see _geometry (default: the uniform fast path for a plain Cartesian mesh, per-point cell
    Jacobians for a mapped one)."""
    if getattr(spec, "is_multifield", False):
        return _flatten_mf(spec, cell_range, with_coef, geometry)
    m = spec.model
    c0, c1 = (0, m.ncells) if cell_range is None else cell_range
    Y, U = spec.Ypde, spec.Ycon
    ny = Y.nloc
    nu = U.nloc if U is not None else 0
    pts, wts = gauss_legendre_01(spec.degree, m.dim)
    nq = pts.shape[0]
    phi_y, dphi_y = lagrange_tables(Y.order, m.dim, pts)
    if U is not None:
        phi_u, _ = lagrange_tables(U.order, m.dim, pts)
    else:
        phi_u = np.zeros((nq, 0))
    ncells = c1 - c0
    cells = np.arange(c0, c1, dtype=np.int64)
    # uniform mesh: physical gradients for the Cartesian map x = lo + h*(idx + xi) folded into the tables
    grad_y, wdet, jinvT, detj, geom_nq = _geometry(m, cells, pts, wts, dphi_y, geometry)

    if with_coef:
        XQ = m.physical_points(cells, pts)
        coef = np.empty((2, ncells, nq))
        for q in range(nq):
            X = XQ[:, q, :]
            coef[0, :, q] = spec.target(X) if spec.target is not None else 0.0
            coef[1, :, q] = spec.source(X) if spec.source is not None else 0.0
    else:
        coef = np.zeros((2, 0, nq))
    params = np.zeros(8)
    params[: len(spec.params)] = spec.params
    return FlatModel(
        spec=spec, cell0=c0, ncells=ncells, dim=m.dim, ny=ny, nu=nu, nq=nq, nparam=spec.nparam,
        nfree_y=Y.nfree, nfree_u=(U.nfree if U is not None else 0),
        cell_dofs_y=np.ascontiguousarray(Y.cell_dof_ids[c0:c1]),
        cell_dofs_u=np.ascontiguousarray(U.cell_dof_ids[c0:c1]) if U is not None else np.zeros((ncells, 0), np.int32),
        dir_y=np.ascontiguousarray(Y.dirichlet_vals, dtype=np.float64),
        dir_u=np.ascontiguousarray(U.dirichlet_vals, dtype=np.float64) if U is not None else np.zeros(0),
        phi_y=np.ascontiguousarray(phi_y), grad_y=np.ascontiguousarray(grad_y),
        phi_u=np.ascontiguousarray(phi_u), wdet=np.ascontiguousarray(wdet),
        coef=np.ascontiguousarray(coef), params=params,
        nfy=1, nfu=(1 if U is not None else 0),
        field_nfree=np.array([Y.nfree] + ([U.nfree] if U is not None else []), dtype=np.int64),
        field_ndir=np.array([Y.ndirichlet] + ([U.ndirichlet] if U is not None else []), dtype=np.int64),
        cell_jinvT=jinvT, cell_detj=detj, geom_nq=geom_nq,
    )


# --------------------------------------------------------------------------
# COO structure, vectorised restatement of the reference constructor's output
# --------------------------------------------------------------------------
def _kept_pairs(row_ids: np.ndarray, col_ids: np.ndarray, lower: bool):
    """Candidates of one block for all cells in the reference loop order
    (col j outer, row i inner).  Returns (rows, cols, keep) shaped
    (ncells, ncols*nrows)."""
    nc, nr = row_ids.shape
    ncol = col_ids.shape[1]
    R = np.broadcast_to(row_ids[:, None, :], (nc, ncol, nr)).reshape(nc, -1)
    C = np.broadcast_to(col_ids[:, :, None], (nc, ncol, nr)).reshape(nc, -1)
    keep = (R > 0) & (C > 0)
    if lower:
        keep &= C <= R
    return R, C, keep


def _field_ids(fm: FlatModel):
    """Per field, the (ncells, nloc) cell dof ids with the multi-field (consecutive) offsets applied to the
    positive ids: state fields first, then control fields (offset by nfree_y)."""
    ys, us = [], []
    off = 0
    for f in range(fm.nfy):
        g = fm.cell_dofs_y[:, f * fm.ny:(f + 1) * fm.ny].astype(np.int64)
        ys.append(np.where(g > 0, g + off, g))
        off += int(fm.field_nfree[f])
    for f in range(fm.nfu):
        g = fm.cell_dofs_u[:, f * fm.nu:(f + 1) * fm.nu].astype(np.int64)
        us.append(np.where(g > 0, g + off, g))
        off += int(fm.field_nfree[fm.nfy + f])
    return ys, us


def _blocks(row_fields, col_fields, lower):
    """Cell-major COO entries of a block matrix: blocks in eachblockid order (column-major over field
    pairs), inside a block column outer / row inner (fill_matrix_hessian_functions.jl:50-94)."""
    Rs, Cs, Ks = [], [], []
    for cid in col_fields:
        for rid in row_fields:
            R, C, keep = _kept_pairs(rid, cid, lower)
            Rs.append(R); Cs.append(C); Ks.append(keep)
    R = np.concatenate(Rs, axis=1); C = np.concatenate(Cs, axis=1); K = np.concatenate(Ks, axis=1)
    return R[K], C[K]


def jac_structure(fm: FlatModel):
    """rows, cols (1-based int64) of jac_nln_structure! for this slab:
    [k-block | y-block | u-block]; /root/reference/src/jacobian_functions.jl:1-7,
    60-120, src/additional_constructors.jl:259-260.  Rows are free dofs of the (multi-field) test space."""
    p = fm.nparam
    ys, us = _field_ids(fm)
    rows = [np.tile(np.arange(1, fm.ncon + 1, dtype=np.int64), p)]
    cols = [np.repeat(np.arange(1, p + 1, dtype=np.int64), fm.ncon)]
    r, c = _blocks(ys, ys, lower=False)
    rows.append(r); cols.append(c + p)
    if us:
        r, c = _blocks(ys, us, lower=False)
        rows.append(r); cols.append(c + p)
    return np.concatenate(rows), np.concatenate(cols)


def _hess_fe_structure(fm: FlatModel):
    ys, us = _field_ids(fm)
    return _blocks(ys + us, ys + us, lower=True)


def _trapezoid(prows: int, p: int):
    rows = [np.arange(j, prows + 1, dtype=np.int64) for j in range(1, p + 1)]
    cols = [np.full(prows - j + 1, j, dtype=np.int64) for j in range(1, p + 1)]
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(rows), np.concatenate(cols)


def hess_structure(fm: FlatModel):
    """rows, cols of hess_structure!: [obj k | obj FE | cons k | cons FE];
    /root/reference/src/hessian_struct_nnzh_functions.jl:10-130,
    src/fill_matrix_hessian_functions.jl:96-209."""
    p = fm.nparam
    nofe = getattr(fm.spec, "no_fe_obj", False)  # NoFETerm: no objective FE block (hessian_struct_nnzh_functions.jl:41-43)
    inde = fm.spec.integrand == KAPPA_POISSON or nofe
    rk, ck = _trapezoid(p if inde else fm.nvar, p)
    rf, cf = _hess_fe_structure(fm)
    rkc, ckc = _trapezoid(fm.nvar, p)
    if nofe:
        return np.concatenate([rk, rkc, rf + p]), np.concatenate([ck, ckc, cf + p]), rk.size
    rows = np.concatenate([rk, rf + p, rkc, rf + p])
    cols = np.concatenate([ck, cf + p, ckc, cf + p])
    return rows, cols, rk.size + rf.size
