"""Bounds of the NLP from functions or vectors — host-side setup, run once per model.

Mirror of ``bounds_functions_to_vectors`` (/root/reference/src/bounds_function.jl:8-214): a bound given
as a function is sampled at the CELL MIDPOINTS and "interpolated" into each field of the space as a
cell-wise constant field, i.e. every free dof takes the value of the last cell that touches it
(``get_free_values(interpolate(cell_l, Yi))``, bounds_function.jl:196-201); a bound given as a vector is
copied.  The result is ``lvar, uvar`` over ``[y free dofs ; u free dofs]`` (the kappa bounds are
prepended by the constructor, src/additional_constructors.jl:190-204).

Functions follow the vectorised convention of problems.py: ``f(X)`` with ``X`` of shape (ncells, dim)
returns (ncells,) for a single-field space or (ncells, nfields).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Union

import numpy as np

from .model_errors import DimensionError

Bound = Union[Callable, np.ndarray, Sequence[float], None]


def _fields(space):
    if space is None:
        return []
    return list(space) if isinstance(space, (list, tuple)) else [space]


def cell_midpoints(model) -> np.ndarray:
    """midpoint(xs) = sum(xs)/length(xs) of the cell vertices (bounds_function.jl:21-24)."""
    if getattr(model, "is_mapped", False):
        return model.vertex_coords()[model.cell_vertices()].mean(axis=1)
    return model.lo + (model.cell_multi_index() + 0.5) * model.h


def _functions_to_vectors(fields, lfunc, ufunc, xm, dtype):
    """_functions_to_vectors! (bounds_function.jl:172-214) for one space."""
    nf = len(fields)
    lo = np.asarray(lfunc(xm), dtype=dtype).reshape(xm.shape[0], -1)
    up = np.asarray(ufunc(xm), dtype=dtype).reshape(xm.shape[0], -1)
    if lo.shape[1] != nf:
        raise DimensionError("lfunc(x)", nf, lo.shape[1])
    if up.shape[1] != nf:
        raise DimensionError("ufunc(x)", nf, up.shape[1])
    ls = [f.interpolate_cellwise_constant(lo[:, i]).astype(dtype) for i, f in enumerate(fields)]
    us = [f.interpolate_cellwise_constant(up[:, i]).astype(dtype) for i, f in enumerate(fields)]
    return np.concatenate(ls) if ls else np.zeros(0, dtype), np.concatenate(us) if us else np.zeros(0, dtype)


def bounds_functions_to_vectors(Ycon, Ypde, model, ly: Bound, uy: Bound, lu: Bound = None, uu: Bound = None,
                                dtype=np.float64):
    """Returns (lvar, uvar) of length num_free_dofs(Ypde) + num_free_dofs(Ycon).

    ``Ypde`` / ``Ycon``: a LagrangianFESpace, a list of them (MultiFieldFESpace) or None (VoidFESpace).
    The five reference methods are the combinations (function | vector) for the state bounds and
    (function | vector | nothing) for the control bounds.
    """
    yf, uf = _fields(Ypde), _fields(Ycon)
    ny, nu = sum(f.nfree for f in yf), sum(f.nfree for f in uf)
    xm = None
    if callable(ly) != callable(uy) or (uf and callable(lu) != callable(uu)):
        raise TypeError("lower and upper bounds of one space must both be functions or both be vectors")
    if callable(ly):
        xm = cell_midpoints(model)
        lvy, uvy = _functions_to_vectors(yf, ly, uy, xm, dtype)
    else:
        lvy, uvy = np.asarray(ly, dtype=dtype), np.asarray(uy, dtype=dtype)
        if lvy.shape != (ny,):
            raise DimensionError("ly", ny, lvy.shape[0] if lvy.ndim == 1 else lvy.shape)
        if uvy.shape != (ny,):
            raise DimensionError("uy", ny, uvy.shape[0] if uvy.ndim == 1 else uvy.shape)
    if not uf:
        return lvy, uvy
    if callable(lu):
        xm = cell_midpoints(model) if xm is None else xm
        lvu, uvu = _functions_to_vectors(uf, lu, uu, xm, dtype)
    else:
        lvu, uvu = np.asarray(lu, dtype=dtype), np.asarray(uu, dtype=dtype)
        if lvu.shape != (nu,):
            raise DimensionError("lu", nu, lvu.shape[0] if lvu.ndim == 1 else lvu.shape)
        if uvu.shape != (nu,):
            raise DimensionError("uu", nu, uvu.shape[0] if uvu.ndim == 1 else uvu.shape)
    return np.concatenate([lvy, lvu]), np.concatenate([uvy, uvu])
