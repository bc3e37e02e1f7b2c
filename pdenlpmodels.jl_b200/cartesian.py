"""Host-side setup that stands in for Gridap 0.15.5 inside this container.

In the real deployment Gridap stays the host-side setup tool (mesh, FE spaces,
dof numbering, Dirichlet data); a Julia layer flattens its objects into the
arrays of ``pdenlp_desc`` (include/pdenlp_cuda.h).  Julia/Gridap are not
available in this image, so this module generates the *same arrays* for the
Cartesian meshes and Lagrangian spaces the BASELINE configs use, following the
Gridap conventions listed in SURVEY.md Appendix A.3:

* ``CartesianDiscreteModel(domain, partition)``: cells and vertices x-fastest,
  reference cell [0,1]^D, cell map x = lo + h*(idx + xi).
* n-cube vertex order x-fastest; QUAD edges: bottom, top, left, right.
* Lagrangian Q_k local dofs grouped by owning face: vertices, edges, interior.
* H1 dof numbering sweeps face dimensions 0..D, faces in global-id order;
  faces carrying a Dirichlet tag get the next negative ids (-1, -2, ...),
  others the next positive ids.  Global edge ids = first encounter looping
  cells then local edges.
* L2 conformity: cell-wise numbering, no sharing.

The numbering is pinned by the reference's own unit tests
(/root/reference/test/unit-test.jl:828,846,855-874), re-expressed in
tests/test_numbering_pins.py.

Everything here is vectorised numpy: the n=2048 membrane mesh (4.2 M cells)
is numbered in a few seconds.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence, Union

import numpy as np

__all__ = [
    "CartesianDiscreteModel",
    "LagrangianFESpace",
    "gauss_legendre_01",
    "lagrange_tables",
]


# --------------------------------------------------------------------------
# mesh
# --------------------------------------------------------------------------
@dataclass
class CartesianDiscreteModel:
    """Mirror of Gridap's ``CartesianDiscreteModel(domain, partition; map=identity)``.

    domain = (lo_1, hi_1, lo_2, hi_2, ...), partition = (n_1, n_2, ...).
    ``map``: optional callable X (npts, dim) -> (npts, dim) applied to the VERTEX coordinates (Gridap's ``map``
    keyword); the cells are then the Q1 (multi-linear) images of the reference cell spanned by their mapped
    vertices, in general neither identical nor affine: the per-cell "weights and Jacobians" of the north star.
    """

    domain: Sequence[float]
    partition: Union[int, Sequence[int]]
    map: Optional[Callable] = None

    def __post_init__(self):
        part = (self.partition,) if np.isscalar(self.partition) else tuple(self.partition)
        self.n = tuple(int(p) for p in part)
        self.dim = len(self.n)
        if len(self.domain) != 2 * self.dim:
            raise ValueError("domain must hold 2*dim numbers")
        self.lo = np.array([float(self.domain[2 * d]) for d in range(self.dim)])
        self.hi = np.array([float(self.domain[2 * d + 1]) for d in range(self.dim)])
        self.h = (self.hi - self.lo) / np.array(self.n, dtype=np.float64)
        self.ncells = int(np.prod(self.n))
        self.nverts = int(np.prod([k + 1 for k in self.n]))

    # cell multi-index, x fastest ------------------------------------------------
    def cell_multi_index(self, cells: Optional[np.ndarray] = None) -> np.ndarray:
        """(ncells, dim) integer multi-index of each cell (0-based)."""
        c = np.arange(self.ncells, dtype=np.int64) if cells is None else np.asarray(cells, dtype=np.int64)
        out = np.empty((c.size, self.dim), dtype=np.int64)
        for d in range(self.dim):
            out[:, d] = c % self.n[d]
            c = c // self.n[d]
        return out

    def cell_vertices(self) -> np.ndarray:
        """(ncells, 2^dim) 0-based vertex ids, n-cube vertex order x-fastest."""
        idx = self.cell_multi_index()
        nv = [k + 1 for k in self.n]
        stride = np.cumprod([1] + nv[:-1])
        base = (idx * stride).sum(axis=1)
        cols = []
        for corner in range(2 ** self.dim):
            off = 0
            for d in range(self.dim):
                if (corner >> d) & 1:
                    off += stride[d]
            cols.append(base + off)
        return np.stack(cols, axis=1)

    def vertex_multi_index(self) -> np.ndarray:
        v = np.arange(self.nverts, dtype=np.int64)
        out = np.empty((v.size, self.dim), dtype=np.int64)
        for d in range(self.dim):
            out[:, d] = v % (self.n[d] + 1)
            v = v // (self.n[d] + 1)
        return out

    @property
    def is_mapped(self) -> bool:
        return self.map is not None

    def vertex_coords(self) -> np.ndarray:
        X = self.lo + self.vertex_multi_index() * self.h
        if self.map is not None:
            X = np.asarray(self.map(X), dtype=np.float64).reshape(X.shape)
        return X

    # geometry of the (possibly mapped) cells ---------------------------------------
    def _cell_vertex_ids(self, cells: np.ndarray) -> np.ndarray:
        idx = self.cell_multi_index(cells)
        nv = [k + 1 for k in self.n]
        stride = np.cumprod([1] + nv[:-1])
        base = (idx * stride).sum(axis=1)
        offs = np.array([sum(stride[d] for d in range(self.dim) if (corner >> d) & 1) for corner in range(2 ** self.dim)])
        return base[:, None] + offs[None, :]

    def physical_points(self, cells: np.ndarray, ref_pts: np.ndarray) -> np.ndarray:
        """(ncells, npts, dim) images of the reference points under each cell map."""
        cells = np.asarray(cells, dtype=np.int64)
        if self.map is None:
            idx = self.cell_multi_index(cells)
            return self.lo + (idx[:, None, :] + ref_pts[None, :, :]) * self.h
        N, _ = lagrange_tables(1, self.dim, ref_pts)          # (npts, 2^dim) Q1 shape functions of the cell map
        xv = self.vertex_coords()[self._cell_vertex_ids(cells)]  # (nc, 2^dim, dim)
        return np.einsum("qv,cvd->cqd", N, xv)

    def cell_jacobians(self, cells: np.ndarray, ref_pts: np.ndarray):
        """J[c, q, d, e] = d x_d / d xi_e of each cell map at the reference points, and |det J|."""
        cells = np.asarray(cells, dtype=np.int64)
        nq = ref_pts.shape[0]
        if self.map is None:
            J = np.broadcast_to(np.diag(self.h), (cells.size, nq, self.dim, self.dim)).copy()
        else:
            _, dN = lagrange_tables(1, self.dim, ref_pts)     # (npts, 2^dim, dim)
            xv = self.vertex_coords()[self._cell_vertex_ids(cells)]
            J = np.einsum("cvd,qve->cqde", xv, dN)
        return J, np.abs(np.linalg.det(J))

    # boundary "entity" of a point given per-direction position flags ------------
    def _entity_from_flags(self, at_lo: np.ndarray, at_hi: np.ndarray) -> np.ndarray:
        """Entity id (1-based) in Gridap's Cartesian face labelling.

        flags are (nfaces, dim) booleans saying whether the face lies on the
        lo / hi side of the box in each direction.  The entity is the index of
        the n-cube face of the bounding box the mesh face lies on: vertices
        first (x-fastest), then edges, ..., interior last.  Implemented for
        dim <= 2 exactly; for dim == 3 boundary faces get entity 1 and interior
        faces 27 (only the tag "boundary" is supported there).
        """
        nb = (at_lo | at_hi).sum(axis=1)
        dim = self.dim
        ent = np.empty(at_lo.shape[0], dtype=np.int64)
        if dim == 1:
            ent[:] = 3
            ent[at_lo[:, 0]] = 1
            ent[at_hi[:, 0]] = 2
        elif dim == 2:
            ent[:] = 9
            # edges of the box: bottom(5) y=lo, top(6) y=hi, left(7) x=lo, right(8) x=hi
            ent[at_lo[:, 1]] = 5
            ent[at_hi[:, 1]] = 6
            ent[at_lo[:, 0]] = 7
            ent[at_hi[:, 0]] = 8
            corner = nb == 2
            cid = 1 + at_hi[:, 0].astype(np.int64) + 2 * at_hi[:, 1].astype(np.int64)
            ent[corner] = cid[corner]
        else:
            ent[:] = 27
            ent[nb > 0] = 1
        return ent

    def nboundary_entities(self) -> int:
        return 3 ** self.dim - 1

    def vertex_entities(self) -> np.ndarray:
        mi = self.vertex_multi_index()
        at_lo = mi == 0
        at_hi = mi == np.array(self.n)
        return self._entity_from_flags(at_lo, at_hi)

    # edges (2D only: needed for Q2) --------------------------------------------
    def cell_edges_2d(self):
        """First-encounter global edge ids for a 2D mesh.

        Returns (cell_edges (ncells,4) 0-based, edge_vertices (nedges,2)).
        Local edge order: bottom (v1,v2), top (v3,v4), left (v1,v3), right (v2,v4).
        """
        assert self.dim == 2
        cv = self.cell_vertices()
        pairs = np.stack(
            [cv[:, [0, 1]], cv[:, [2, 3]], cv[:, [0, 2]], cv[:, [1, 3]]], axis=1
        )  # (nc, 4, 2), already sorted since vertex ids grow with x then y
        keys = pairs[:, :, 0] * np.int64(self.nverts) + pairs[:, :, 1]
        flat = keys.reshape(-1)
        uniq, first, inv = np.unique(flat, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")  # unique-key index -> encounter order
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        cell_edges = rank[inv].reshape(self.ncells, 4)
        ev = np.stack([uniq[order] // self.nverts, uniq[order] % self.nverts], axis=1)
        return cell_edges, ev

    def edge_entities_2d(self, edge_vertices: np.ndarray) -> np.ndarray:
        vmi = self.vertex_multi_index()
        a = vmi[edge_vertices[:, 0]]
        b = vmi[edge_vertices[:, 1]]
        n = np.array(self.n)
        at_lo = (a == 0) & (b == 0)
        at_hi = (a == n) & (b == n)
        return self._entity_from_flags(at_lo, at_hi)


# --------------------------------------------------------------------------
# reference FE tables and quadrature
# --------------------------------------------------------------------------
def gauss_legendre_01(degree: int, dim: int):
    """Tensor Gauss-Legendre on [0,1]^dim with ceil((degree+1)/2) points per
    direction, first direction fastest (Gridap ``Measure(trian, degree)`` on
    n-cubes).  Returns (points (nq, dim), weights (nq,))."""
    npd = int(math.ceil((degree + 1) / 2))
    t, w = np.polynomial.legendre.leggauss(npd)
    xi1 = 0.5 * (t + 1.0)
    w1 = 0.5 * w
    nq = npd ** dim
    pts = np.empty((nq, dim))
    wts = np.ones(nq)
    for q in range(nq):
        r = q
        for d in range(dim):
            k = r % npd
            r //= npd
            pts[q, d] = xi1[k]
            wts[q] *= w1[k]
    return pts, wts


def _lagrange_1d(order: int, xi: np.ndarray):
    """Values and derivatives of the 1-D nodal Lagrange basis on [0,1].
    Node labels: 0 -> xi=0, 1 -> xi=1, 2 -> xi=1/2 (order 2 only)."""
    xi = np.asarray(xi, dtype=np.float64)
    if order == 1:
        val = np.stack([1.0 - xi, xi], axis=0)
        der = np.stack([-np.ones_like(xi), np.ones_like(xi)], axis=0)
    elif order == 2:
        val = np.stack([(1.0 - xi) * (1.0 - 2.0 * xi), xi * (2.0 * xi - 1.0), 4.0 * xi * (1.0 - xi)], axis=0)
        der = np.stack([4.0 * xi - 3.0, 4.0 * xi - 1.0, 4.0 - 8.0 * xi], axis=0)
    else:
        raise NotImplementedError("Lagrangian order must be 1 or 2")
    return val, der


def local_node_labels(order: int, dim: int):
    """Per local dof, the 1-D node label in each direction, in Gridap's local
    dof order (vertices, edges, [faces], interior)."""
    if order == 1:
        return [tuple((c >> d) & 1 for d in range(dim)) for c in range(2 ** dim)]
    if order == 2 and dim == 1:
        return [(0,), (1,), (2,)]
    if order == 2 and dim == 2:
        verts = [(0, 0), (1, 0), (0, 1), (1, 1)]
        edges = [(2, 0), (2, 1), (0, 2), (1, 2)]  # bottom, top, left, right
        return verts + edges + [(2, 2)]
    raise NotImplementedError("Q2 is implemented for dim <= 2")


def lagrange_tables(order: int, dim: int, pts: np.ndarray):
    """Shape values (nq, nloc) and reference gradients (nq, nloc, dim)."""
    labels = local_node_labels(order, dim)
    nq = pts.shape[0]
    nloc = len(labels)
    vals1 = []
    ders1 = []
    for d in range(dim):
        v, g = _lagrange_1d(order, pts[:, d])
        vals1.append(v)
        ders1.append(g)
    phi = np.ones((nq, nloc))
    grad = np.ones((nq, nloc, dim))
    for a, lab in enumerate(labels):
        for d in range(dim):
            phi[:, a] *= vals1[d][lab[d]]
            for e in range(dim):
                grad[:, a, e] *= ders1[d][lab[d]] if d == e else vals1[d][lab[d]]
    return phi, grad


def local_node_coords(order: int, dim: int) -> np.ndarray:
    pos = {0: 0.0, 1: 1.0, 2: 0.5}
    return np.array([[pos[l] for l in lab] for lab in local_node_labels(order, dim)])


# --------------------------------------------------------------------------
# FE space
# --------------------------------------------------------------------------
DirichletSpec = Union[None, str, Sequence[int]]


@dataclass
class LagrangianFESpace:
    """Scalar Lagrangian Q_k space on a Cartesian model.

    Mirrors ``TestFESpace(model, ReferenceFE(lagrangian, Float64, order);
    conformity, dirichlet_tags)`` + ``TrialFESpace(V, g)``.

    dirichlet: None | "boundary" | list of entity ids (e.g. [1, 2] in 1-D for
    the tags built with ``add_tag_from_tags!(labels, "diri0", [1])``).
    dirichlet_values: callable g(x)->float applied at Dirichlet nodes, a float,
    or a dict {entity id: value} (per-tag values, burger1d.jl:26-28).
    """

    model: CartesianDiscreteModel
    order: int = 1
    conformity: str = "H1"
    dirichlet: DirichletSpec = None
    dirichlet_values: Union[None, float, dict, Callable] = None

    cell_dof_ids: np.ndarray = field(init=False, repr=False)
    nfree: int = field(init=False)
    ndirichlet: int = field(init=False)
    dirichlet_vals: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        m = self.model
        self.dim = m.dim
        self.nloc = (self.order + 1) ** m.dim
        if self.conformity.upper() == "L2":
            ids = np.arange(1, m.ncells * self.nloc + 1, dtype=np.int64).reshape(m.ncells, self.nloc)
            self.cell_dof_ids = ids.astype(np.int32)
            self.nfree = m.ncells * self.nloc
            self.ndirichlet = 0
            self.dirichlet_vals = np.zeros(0)
            return
        if self.conformity.upper() != "H1":
            raise ValueError("conformity must be H1 or L2")
        self._number_h1()

    # -- H1 numbering --------------------------------------------------------
    def _is_dirichlet_entity(self, ent: np.ndarray) -> np.ndarray:
        if self.dirichlet is None:
            return np.zeros(ent.shape, dtype=bool)
        nb = self.model.nboundary_entities()
        if isinstance(self.dirichlet, str):
            if self.dirichlet != "boundary":
                raise ValueError("only the tag 'boundary' is known by name")
            return ent <= nb
        if self.model.dim == 3:
            raise NotImplementedError("entity lists are implemented for dim <= 2")
        return np.isin(ent, np.asarray(list(self.dirichlet), dtype=np.int64))

    def _dir_value_at(self, coords: np.ndarray, ent: np.ndarray) -> np.ndarray:
        g = self.dirichlet_values
        if g is None:
            return np.zeros(coords.shape[0])
        if isinstance(g, dict):
            out = np.zeros(coords.shape[0])
            for e, v in g.items():
                out[ent == e] = float(v)
            return out
        if callable(g):
            return np.asarray(g(coords), dtype=np.float64) * np.ones(coords.shape[0])
        return np.full(coords.shape[0], float(g))

    def _number_h1(self):
        m = self.model
        nfree = 0
        ndir = 0
        dir_vals = []
        blocks = []  # per-cell id columns in local dof order

        def sweep(is_dir, coords, ent):
            nonlocal nfree, ndir
            ids = np.empty(is_dir.shape[0], dtype=np.int64)
            nd = np.cumsum(is_dir)
            nf = np.cumsum(~is_dir)
            ids[is_dir] = -(ndir + nd[is_dir])
            ids[~is_dir] = nfree + nf[~is_dir]
            if is_dir.any():
                dir_vals.append(self._dir_value_at(coords[is_dir], ent[is_dir]))
            ndir += int(nd[-1]) if nd.size else 0
            nfree += int(nf[-1]) if nf.size else 0
            return ids

        # d = 0: vertices (one dof each for order >= 1)
        vent = m.vertex_entities()
        vids = sweep(self._is_dirichlet_entity(vent), m.vertex_coords(), vent)
        blocks.append(vids[m.cell_vertices()])

        if self.order == 2:
            if m.dim == 1:
                # d = 1 is the cell itself: interior node, never Dirichlet
                cent = np.full(m.ncells, 3, dtype=np.int64)
                mids = m.lo + (m.cell_multi_index() + 0.5) * m.h
                cids = sweep(np.zeros(m.ncells, dtype=bool), mids, cent)
                blocks.append(cids[:, None])
            elif m.dim == 2:
                ce, ev = m.cell_edges_2d()
                eent = m.edge_entities_2d(ev)
                vc = m.vertex_coords()
                emid = 0.5 * (vc[ev[:, 0]] + vc[ev[:, 1]])
                eids = sweep(self._is_dirichlet_entity(eent), emid, eent)
                blocks.append(eids[ce])
                cent = np.full(m.ncells, 9, dtype=np.int64)
                mids = m.lo + (m.cell_multi_index() + 0.5) * m.h
                cids = sweep(np.zeros(m.ncells, dtype=bool), mids, cent)
                blocks.append(cids[:, None])
            else:
                raise NotImplementedError("Q2 is implemented for dim <= 2")
        elif self.order != 1:
            raise NotImplementedError("Lagrangian order must be 1 or 2")

        ids = np.concatenate(blocks, axis=1)
        assert ids.shape == (m.ncells, self.nloc)
        if max(nfree, ndir) >= 2 ** 31:
            raise OverflowError("dof ids exceed int32")
        self.cell_dof_ids = ids.astype(np.int32)
        self.nfree = nfree
        self.ndirichlet = ndir
        self.dirichlet_vals = np.concatenate(dir_vals) if dir_vals else np.zeros(0)

    # -- helpers -------------------------------------------------------------
    def num_free_dofs(self) -> int:
        return self.nfree

    def interpolate_cellwise_constant(self, cell_values: np.ndarray) -> np.ndarray:
        """Free values of ``interpolate(cell_l, V)`` for a cell-wise constant
        field: the last cell touching a dof wins (bounds.py and the numbering
        pins; /root/reference/src/bounds_function.jl:196-201)."""
        out = np.zeros(self.nfree)
        ids = self.cell_dof_ids
        vals = np.broadcast_to(np.asarray(cell_values, dtype=np.float64)[:, None], ids.shape)
        free = ids > 0
        # cell-major order; for a repeated index NumPy keeps the LAST assigned value = the last cell
        out[ids[free] - 1] = vals[free]
        return out
