"""Cell-slab sharding of one model over the GPUs of a box (one process per GPU).

SURVEY §8(e): cells are independent units and are numbered x-fastest, so a
contiguous cell range is a slab of mesh rows/planes.  Rank r owns cells
[c_r, c_{r+1}) and therefore ONE contiguous sub-range of every COO segment:
``jac_coord!`` / ``hess_coord!`` need no communication — each rank writes its own
slice, the concatenation over ranks (segment by segment) is the global stream.

The dof-indexed outputs (``obj``, ``grad!``, ``cons!``, ``jprod!``, ``jtprod!``,
``hprod!``) are partial sums per slab; only the *interface* dofs (touched by more
than one slab: one mesh row of state dofs + one row of control dofs per cut) and
the kappa entries need a reduction: they are packed into one small buffer and
all-reduced (NCCL over NVLink on the GPUs, gloo in the CPU tests).  After the
reduction every rank holds the exact value of every dof its slab touches and
zero elsewhere; ``allgather_owned`` assembles the replicated vector when a
caller wants one.

``local`` is any object with the GridapPDENLPModel callback interface: the CUDA
model in production, an oracle-backed stand-in in the world_size-2 gloo tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def partition_cells(ncells: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Contiguous slabs with boundaries aligned to the 32-cell tiles of the kernels."""
    ntiles = (ncells + align - 1) // align
    out = []
    for r in range(world):
        t0, t1 = ntiles * r // world, ntiles * (r + 1) // world
        out.append((min(t0 * align, ncells), min(t1 * align, ncells)))
    return out


def _touched(ids: np.ndarray, nfree: int, ranges: Sequence[Tuple[int, int]]) -> np.ndarray:
    """(world, nfree) boolean: dof touched by slab r."""
    out = np.zeros((len(ranges), nfree), dtype=bool)
    for r, (c0, c1) in enumerate(ranges):
        g = ids[c0:c1].reshape(-1)
        g = g[g > 0]
        out[r, g - 1] = True
    return out


def _space_ids(fields) -> Tuple[np.ndarray, int]:
    """Cell dof ids of a (multi-field) space with the positive ids offset to the concatenated numbering."""
    cols, off = [], 0
    for f in fields:
        g = f.cell_dof_ids.astype(np.int64)
        cols.append(np.where(g > 0, g + off, g))
        off += f.nfree
    return (np.concatenate(cols, axis=1) if cols else np.zeros((0, 0), np.int64)), off


class SlabLayout:
    """Index sets of the slab decomposition, in x-layout (0-based, kappa first)."""

    def __init__(self, spec, world: int):
        self.world = world
        self.ranges = partition_cells(spec.model.ncells, world)
        p = spec.nparam
        if getattr(spec, "is_multifield", False):
            yfields, ufields = list(spec.Yfields), list(spec.Ufields)
        else:
            yfields, ufields = [spec.Ypde], ([spec.Ycon] if spec.Ycon is not None else [])
        ids_y, ny = _space_ids(yfields)
        ids_u, nu = _space_ids(ufields)
        ty = _touched(ids_y, ny, self.ranges)
        self.touch_y = ty
        iface_y = np.nonzero(ty.sum(axis=0) > 1)[0]
        if nu:
            tu = _touched(ids_u, nu, self.ranges)
            iface_u = np.nonzero(tu.sum(axis=0) > 1)[0]
        else:
            tu = np.zeros((world, 0), dtype=bool)
            iface_u = np.zeros(0, dtype=np.int64)
        self.touch_u = tu
        # interface entries of an nvar-vector (kappa entries are shared by every slab)
        self.iface_var = np.concatenate([np.arange(p), p + iface_y, p + ny + iface_u]).astype(np.int64)
        # interface entries of an ncon-vector (rows = free dofs of the test space Xpde = Ypde ids)
        self.iface_con = iface_y.astype(np.int64)
        # ownership for allgather_owned: the lowest rank touching a dof owns it
        self.owner_y = np.argmax(ty, axis=0)
        self.owner_u = np.argmax(tu, axis=0) if nu else np.zeros(0, dtype=np.int64)
        self.p, self.ny, self.nu = p, ny, nu

    # ---- dense kappa blocks (dof-indexed like the vector outputs: only interface rows need a reduction) ----------
    def trap_offsets(self, prows: int) -> np.ndarray:
        """Start of column j (0-based) of a lower trapezoid with `prows` rows and p columns, "i fastest", j <= i
        (src/additional_obj_terms.jl:311, src/GridapPDENLPModel.jl:362)."""
        j = np.arange(self.p + 1, dtype=np.int64)
        return j * prows - j * (j - 1) // 2

    def jac_kblock_index(self, rows: np.ndarray) -> np.ndarray:
        """Positions of the given constraint rows (0-based) in the column-major ncon x p Jacobian kappa-block."""
        return (np.arange(self.p, dtype=np.int64)[:, None] * self.ny + rows[None, :]).reshape(-1)

    def trap_index(self, prows: int, var_rows: np.ndarray) -> np.ndarray:
        """Positions of the given x-layout rows (0-based, kappa rows included) in a Hessian kappa trapezoid."""
        off = self.trap_offsets(prows)
        out = []
        for j in range(self.p):
            r = var_rows[(var_rows >= j) & (var_rows < prows)]
            out.append(off[j] + (r - j))
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)

    def var_owner(self) -> np.ndarray:
        """Owning rank of every x-layout entry (kappa entries: rank 0)."""
        return np.concatenate([np.zeros(self.p, dtype=np.int64), self.owner_y, self.owner_u]).astype(np.int64)


class ShardedModel:
    def __init__(self, spec, local, layout: SlabLayout, rank: int, group=None):
        self.spec, self.local, self.layout, self.rank, self.group = spec, local, layout, rank, group
        self.world = layout.world
        self._idx_cache = {}
        # NoFETerm objective fk(kappa): not a sum over cells, so only ONE slab (rank 0) may contribute it to the
        # reduced outputs (every slab's model evaluates the whole of it)
        self.nofe = bool(getattr(spec, "no_fe_obj", False))

    # ---- helpers ---------------------------------------------------------
    def _idx(self, which: str, like):
        key = (which, like.device if isinstance(like, torch.Tensor) else "np")
        if key not in self._idx_cache:
            idx = self.layout.iface_var if which == "var" else self.layout.iface_con
            t = torch.from_numpy(idx)
            if isinstance(like, torch.Tensor):
                t = t.to(like.device)
            self._idx_cache[key] = t
        return self._idx_cache[key]

    def _reduce_interface(self, vec, which: str):
        """Sum the interface entries of a slab-partial vector over the ranks."""
        if self.world == 1:
            return vec
        t = vec if isinstance(vec, torch.Tensor) else torch.from_numpy(vec)
        idx = self._idx(which, t)
        if idx.numel():
            buf = t.index_select(0, idx).contiguous()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            t.index_copy_(0, idx, buf)
        return vec

    # ---- scalar ----------------------------------------------------------
    def obj(self, x) -> float:
        f = self.local.obj(x)
        if self.world == 1 or self.nofe:
            return f
        dev = x.device if isinstance(x, torch.Tensor) else "cpu"
        t = torch.tensor([f], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    # ---- dof-indexed outputs --------------------------------------------
    def grad_(self, x, g):
        g = self.local.grad_(x, g)
        if self.nofe and self.rank != 0:
            g[: self.layout.p] = 0.0
        return self._reduce_interface(g, "var")

    def cons_(self, x, c):
        return self._reduce_interface(self.local.cons_(x, c), "con")

    def jprod_(self, x, v, Jv):
        return self._reduce_interface(self.local.jprod_(x, v, Jv), "con")

    def jtprod_(self, x, v, Jtv):
        return self._reduce_interface(self.local.jtprod_(x, v, Jtv), "var")

    def _ow(self, obj_weight):
        return 0.0 if (self.nofe and self.rank != 0) else obj_weight

    def hprod_(self, x, *args, obj_weight: float = 1.0):
        return self._reduce_interface(self.local.hprod_(x, *args, obj_weight=self._ow(obj_weight)), "var")

    # ---- COO streams: local slices of the FE segments, no communication.  The dense kappa blocks are dof-indexed
    #      (not cell-indexed) sums over cells, i.e. slab partials exactly like grad!/cons!: their INTERFACE rows (and
    #      the p x p kappa-kappa corner) are packed and all-reduced; afterwards every rank holds the exact value of
    #      every row its slab touches and zero elsewhere ------------------------------------------------------------
    def _reduce_packed(self, vals, base: int, idx_fn, key):
        """Sum the entries base + idx_fn() of a COO value vector over the ranks (one packed all-reduce).  The index set
        is built once per (key, device): it is a function of the slab layout only (rebuilding it with numpy on every
        callback cost 3 ms per hess_coord! on config 5 at 8 GPUs)."""
        if self.world == 1:
            return vals
        t = vals if isinstance(vals, torch.Tensor) else torch.from_numpy(vals)
        # (several blocks in one exchange: base / idx_fn are tuples, the index sets are concatenated)
        bases = base if isinstance(base, tuple) else (base,)
        fns = idx_fn if isinstance(idx_fn, tuple) else (idx_fn,)
        ck = (key, t.device, bases)
        if ck not in self._idx_cache:
            parts = [np.asarray(f(), dtype=np.int64) + b for f, b in zip(fns, bases)]
            self._idx_cache[ck] = torch.from_numpy(np.concatenate(parts) if parts else np.zeros(0, np.int64)).to(t.device)
        ix = self._idx_cache[ck]
        if ix.numel() == 0:
            return vals
        buf = t.index_select(0, ix).contiguous()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        t.index_copy_(0, ix, buf)
        return vals

    def _prows_obj(self) -> int:
        pm = self.local.pdemeta
        p = self.layout.p
        return p if pm.nnzh_k_obj == p * (p + 1) // 2 else p + self.layout.ny + self.layout.nu

    def jac_coord_(self, x, vals):
        """[k-block (ncon x p) | this slab's y-block | this slab's u-block]; interface rows of the kappa-block reduced
        (src/jacobian_functions.jl:135-142)."""
        vals = self.local.jac_coord_(x, vals)
        if self.layout.p == 0:
            return vals
        return self._reduce_packed(vals, 0, lambda: self.layout.jac_kblock_index(self.layout.iface_con), "jk")

    def hess_coord_(self, x, *args, obj_weight: float = 1.0):
        """[obj k-block | this slab's obj FE | cons k-block | this slab's cons FE]; interface rows of both kappa
        trapezoids reduced (src/GridapPDENLPModel.jl:361-380, src/additional_obj_terms.jl:286-314)."""
        vals = self.local.hess_coord_(x, *args, obj_weight=self._ow(obj_weight))
        L, pm = self.layout, self.local.pdemeta
        if L.p == 0:
            return vals
        nvar = L.p + L.ny + L.nu
        # one packed all-reduce for both trapezoids (two exchanges cost 0.1 ms of launch + NCCL latency at 8 GPUs)
        return self._reduce_packed(vals, (0, int(pm.nnzh_obj)),
                                   (lambda: L.trap_index(self._prows_obj(), L.iface_var), lambda: L.trap_index(nvar, L.iface_var)), "hk")

    def assemble_global_coo(self, vals, which: str):
        """The GLOBAL COO value stream (replicated on every rank) from the per-rank streams: every dense kappa-block
        row from its owner, then each FE segment as the concatenation of the slab slices in rank order.
        Verification helper (O(nnz) traffic); which = "jac" | "hess"."""
        pm, L = self.local.pdemeta, self.layout
        nvar = L.p + L.ny + L.nu
        t = (vals if isinstance(vals, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vals))).clone()
        own_var = np.nonzero(L.var_owner() == self.rank)[0]
        own_con = np.nonzero(L.owner_y == self.rank)[0]
        if which == "jac":
            dense = [((0, pm.nnzj_k), L.jac_kblock_index(own_con))]
            fe = [(pm.nnzj_k, pm.nnzj_k + pm.nnzj_y), (pm.nnzj_k + pm.nnzj_y, pm.nnzj_k + pm.nnzj_y + pm.nnzj_u)]
            order = [("d", 0), ("f", 0), ("f", 1)]
        else:
            nfo = pm.nnzh_obj - pm.nnzh_k_obj
            dense = [((0, pm.nnzh_k_obj), L.trap_index(self._prows_obj(), own_var)),
                     ((pm.nnzh_obj, pm.nnzh_obj + pm.nnzh_k_con), L.trap_index(nvar, own_var))]
            fe = [(pm.nnzh_k_obj, pm.nnzh_k_obj + nfo),
                  (pm.nnzh_obj + pm.nnzh_k_con, pm.nnzh_obj + pm.nnzh_k_con + pm.nnzh_fe)]
            order = [("d", 0), ("f", 0), ("d", 1), ("f", 1)]
        dense_out = []
        for (a, b), own in dense:
            blk = t[a:b].clone()
            if self.world > 1 and b > a:
                keep = torch.zeros(b - a, dtype=torch.bool, device=blk.device)
                keep[torch.from_numpy(own).to(blk.device)] = True
                blk = torch.where(keep, blk, torch.zeros_like(blk))
                dist.all_reduce(blk, op=dist.ReduceOp.SUM, group=self.group)
            dense_out.append(blk)
        gathered = []
        for a, b in fe:
            mine = t[a:b].contiguous()
            if self.world == 1:
                gathered.append(mine)
                continue
            n = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
            sizes = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(sizes, n, group=self.group)
            sizes = [int(v.item()) for v in sizes]
            pad = torch.zeros(max(sizes + [1]), dtype=mine.dtype, device=mine.device)
            pad[: mine.numel()] = mine
            parts = [torch.zeros_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=self.group)
            gathered.append(torch.cat([q[:k] for q, k in zip(parts, sizes)]))
        return torch.cat([dense_out[i] if kind == "d" else gathered[i] for kind, i in order])

    def segment_sizes(self) -> dict:
        pm = self.local.pdemeta
        return {"jy": pm.nnzj_y, "ju": pm.nnzj_u, "hfe": pm.nnzh_fe}

    def global_segment_offsets(self) -> dict:
        """Offset of this rank's slice inside each global FE segment (exclusive scan over ranks)."""
        sizes = self.segment_sizes()
        keys = sorted(sizes)
        mine = torch.tensor([sizes[k] for k in keys], dtype=torch.int64)
        if self.world == 1:
            return {k: 0 for k in keys}
        dev = self.local.device if hasattr(self.local, "device") and dist.get_backend(self.group) == "nccl" else "cpu"
        mine = mine.to(dev)
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine, group=self.group)
        allv = torch.stack(allv).cpu().numpy()
        off = allv[: self.rank].sum(axis=0) if self.rank else np.zeros(len(keys), dtype=np.int64)
        return {k: int(o) for k, o in zip(keys, off)}

    def allgather_owned(self, vec, which: str = "var"):
        """Replicated full vector from interface-reduced slab vectors: every dof is taken from
        its owner (the lowest rank touching it).  Debug/verification helper, O(n) traffic."""
        if self.world == 1:
            return vec
        L = self.layout
        t = (vec if isinstance(vec, torch.Tensor) else torch.from_numpy(vec)).clone()
        if which == "var":
            own = np.concatenate([np.zeros(L.p, dtype=np.int64), L.owner_y, L.owner_u])
        else:
            own = L.owner_y
        mask = torch.from_numpy(own == self.rank).to(t.device)
        t = torch.where(mask, t, torch.zeros_like(t))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t
