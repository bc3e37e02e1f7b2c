"""ctypes binding of include/pdenlp_cuda.h (libpdenlp_cuda.so).

This is the Python counterpart of the Julia ``ccall`` layer shown in
INTEGRATION.md / julia/PDENLPCuda.jl: plain pointers and sizes only.
There is no fallback: if the library is missing or no sm_100 device is
present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PDENLP_CUDA_LIB selects an alternative build of the same library (A/B kernel experiments)
LIB_PATH = os.environ.get("PDENLP_CUDA_LIB") or os.path.join(HERE, "csrc", "libpdenlp_cuda.so")

ABI_VERSION = 3
OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM = 0, -1, -2, -3, -4
MEM_DEVICE, MEM_HOST = 0, 1
FLAG_DETERMINISTIC = 1
FLAG_NO_TEMPLATES = 2

# every symbol include/pdenlp_cuda.h declares
SYMBOLS = [
    "pdenlp_last_error", "pdenlp_abi_version", "pdenlp_build_id", "pdenlp_create", "pdenlp_destroy", "pdenlp_get_info",
    "pdenlp_set_stream", "pdenlp_sync", "pdenlp_malloc", "pdenlp_free", "pdenlp_memcpy_h2d", "pdenlp_memcpy_d2h",
    "pdenlp_host_alloc", "pdenlp_host_free",
    "pdenlp_obj", "pdenlp_objgrad", "pdenlp_objcons", "pdenlp_grad", "pdenlp_cons", "pdenlp_jac_coord", "pdenlp_hess_coord", "pdenlp_hess_coord_obj",
    "pdenlp_jprod", "pdenlp_jtprod", "pdenlp_hprod", "pdenlp_jac_structure", "pdenlp_hess_structure", "pdenlp_input_ranges",
    "pdenlp_sharded_create", "pdenlp_sharded_destroy", "pdenlp_sharded_get_info", "pdenlp_sharded_slab", "pdenlp_sharded_obj", "pdenlp_sharded_grad", "pdenlp_sharded_cons", "pdenlp_sharded_jac_coord", "pdenlp_sharded_hess_coord", "pdenlp_sharded_jprod", "pdenlp_sharded_jtprod", "pdenlp_sharded_hprod", "pdenlp_sharded_jac_structure", "pdenlp_sharded_hess_structure",
    "pdenlp_selfcheck_error", "pdenlp_launch_count", "pdenlp_template_stats", "pdenlp_graph_replays", "pdenlp_last_kernel_ms", "pdenlp_set_timing",
    "pdenlp_compaction_create", "pdenlp_compaction_destroy", "pdenlp_compaction_nnz",
    "pdenlp_compaction_structure", "pdenlp_compaction_apply",
]
ORDER_CSC, ORDER_CSR = 0, 1


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("integrand", C.c_int32), ("dim", C.c_int32), ("ny", C.c_int32),
        ("nu", C.c_int32), ("nq", C.c_int32), ("nparam", C.c_int32), ("memspace", C.c_int32),
        ("device", C.c_int32), ("nfields_y", C.c_int32),
        ("ncells", C.c_int64), ("nfree_y", C.c_int64), ("nfree_u", C.c_int64), ("ndir_y", C.c_int64), ("ndir_u", C.c_int64),
        ("cell_dofs_y", C.c_void_p), ("cell_dofs_u", C.c_void_p), ("dir_y", C.c_void_p), ("dir_u", C.c_void_p),
        ("phi_y", C.c_void_p), ("grad_y", C.c_void_p), ("phi_u", C.c_void_p), ("wdet", C.c_void_p),
        ("ncoef", C.c_int32), ("nfields_u", C.c_int32), ("coef", C.c_void_p),
        ("params", C.c_double * 8),
        ("field_nfree", C.c_void_p), ("field_ndir", C.c_void_p),
        ("cell_jinvT", C.c_void_p), ("cell_detj", C.c_void_p), ("geom_nq", C.c_int32), ("flags", C.c_int32),
    ]


class Info(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "nvar", "ncon", "nnzj", "nnzj_k", "nnzj_y", "nnzj_u", "nnzh", "nnzh_obj", "nnzh_k_obj", "nnzh_fe", "nnzh_k_con",
        "nnzh_fe_obj")]


class PdenlpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpdenlp_cuda error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libpdenlp_cuda.so and check that it exports every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is not built; run `python __graft_entry__.py` (nvcc, sm_100a). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"libpdenlp_cuda.so lacks symbols: {missing}")
    lib.pdenlp_last_error.restype = C.c_char_p
    lib.pdenlp_build_id.restype = C.c_char_p
    # build provenance: the binary must have been built from the sources of this tree (PDENLP_CUDA_LIB selects an
    # experimental build on purpose and is exempt)
    if not os.environ.get("PDENLP_CUDA_LIB"):
        from .build import source_hash
        have, want = lib.pdenlp_build_id().decode(), source_hash()
        if have != want:
            raise ImportError(f"{LIB_PATH} is stale: built from sources {have}, the tree is {want}; "
                              "run `python __graft_entry__.py` to rebuild (there is no fallback path)")
    vp, dbl, i64 = C.c_void_p, C.c_double, C.c_int64
    lib.pdenlp_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    lib.pdenlp_destroy.argtypes = [vp]
    lib.pdenlp_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.pdenlp_set_stream.argtypes = [vp, vp]
    lib.pdenlp_sync.argtypes = [vp]
    lib.pdenlp_malloc.argtypes = [C.POINTER(vp), i64]
    lib.pdenlp_free.argtypes = [vp]
    lib.pdenlp_memcpy_h2d.argtypes = [vp, vp, i64]
    lib.pdenlp_memcpy_d2h.argtypes = [vp, vp, i64]
    lib.pdenlp_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.pdenlp_host_free.argtypes = [vp]
    lib.pdenlp_obj.argtypes = [vp, vp, C.POINTER(dbl)]
    lib.pdenlp_grad.argtypes = [vp, vp, vp]
    lib.pdenlp_objgrad.argtypes = [vp, vp, C.POINTER(dbl), vp]
    lib.pdenlp_objcons.argtypes = [vp, vp, C.POINTER(dbl), vp]
    lib.pdenlp_cons.argtypes = [vp, vp, vp]
    lib.pdenlp_jac_coord.argtypes = [vp, vp, vp]
    lib.pdenlp_hess_coord.argtypes = [vp, vp, vp, dbl, vp]
    lib.pdenlp_hess_coord_obj.argtypes = [vp, vp, dbl, vp]
    lib.pdenlp_jprod.argtypes = [vp, vp, vp, vp]
    lib.pdenlp_jtprod.argtypes = [vp, vp, vp, vp]
    lib.pdenlp_hprod.argtypes = [vp, vp, vp, dbl, vp, vp]
    lib.pdenlp_jac_structure.argtypes = [vp, vp, vp]
    lib.pdenlp_hess_structure.argtypes = [vp, vp, vp]
    lib.pdenlp_selfcheck_error.argtypes = [vp, C.POINTER(dbl)]
    lib.pdenlp_input_ranges.argtypes = [vp, C.POINTER(i64)]
    lib.pdenlp_sharded_create.argtypes = [C.POINTER(Desc), C.c_int32, C.POINTER(C.c_int32), C.POINTER(vp)]
    lib.pdenlp_sharded_destroy.argtypes = [vp]
    lib.pdenlp_sharded_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.pdenlp_sharded_slab.argtypes = [vp, C.c_int32, C.POINTER(i64), C.POINTER(i64)]
    lib.pdenlp_sharded_obj.argtypes = [vp, vp, C.POINTER(dbl)]
    lib.pdenlp_sharded_grad.argtypes = [vp, vp, vp]
    lib.pdenlp_sharded_cons.argtypes = [vp, vp, vp]
    lib.pdenlp_sharded_jac_coord.argtypes = [vp, vp, vp]
    lib.pdenlp_sharded_hess_coord.argtypes = [vp, vp, vp, dbl, vp]
    lib.pdenlp_sharded_jprod.argtypes = [vp, vp, vp, vp]
    lib.pdenlp_sharded_jtprod.argtypes = [vp, vp, vp, vp]
    lib.pdenlp_sharded_hprod.argtypes = [vp, vp, vp, dbl, vp, vp]
    lib.pdenlp_sharded_jac_structure.argtypes = [vp, vp, vp]
    lib.pdenlp_sharded_hess_structure.argtypes = [vp, vp, vp]
    lib.pdenlp_launch_count.argtypes = [vp, C.POINTER(i64)]
    lib.pdenlp_template_stats.argtypes = [vp, C.POINTER(i64)]
    lib.pdenlp_graph_replays.argtypes = [vp, C.POINTER(i64)]
    lib.pdenlp_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.pdenlp_set_timing.argtypes = [vp, C.c_int32]
    lib.pdenlp_compaction_create.argtypes = [i64, vp, vp, i64, i64, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.pdenlp_compaction_destroy.argtypes = [vp]
    lib.pdenlp_compaction_nnz.argtypes = [vp, C.POINTER(i64)]
    lib.pdenlp_compaction_structure.argtypes = [vp, vp, vp, vp, C.c_int32]
    lib.pdenlp_compaction_apply.argtypes = [vp, vp, vp, vp]
    _lib = lib
    return lib


def build_id() -> str:
    return load().pdenlp_build_id().decode()


def check(rc):
    if rc != OK:
        raise PdenlpError(rc, load().pdenlp_last_error().decode())


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def build_desc(fm, device: int = -1, memspace: int = MEM_DEVICE, with_coef: bool = True, flags: int = 0):
    """pdenlp_desc of a FlatModel plus the arrays it points at (keep them alive until pdenlp_create returns)."""
    d = Desc()
    d.abi_version = ABI_VERSION
    d.integrand = fm.spec.integrand
    d.dim, d.ny, d.nu, d.nq, d.nparam = fm.dim, fm.ny, fm.nu, fm.nq, fm.nparam
    d.memspace = memspace
    d.device = device
    d.flags = flags
    d.ncells, d.nfree_y, d.nfree_u = fm.ncells, fm.nfree_y, fm.nfree_u
    d.ndir_y, d.ndir_u = fm.dir_y.size, fm.dir_u.size
    keep = [np.ascontiguousarray(fm.cell_dofs_y, np.int32), np.ascontiguousarray(fm.cell_dofs_u, np.int32),
            fm.dir_y, fm.dir_u, fm.phi_y, fm.grad_y, fm.phi_u, fm.wdet, fm.coef]
    d.cell_dofs_y, d.cell_dofs_u = _np_ptr(keep[0]), _np_ptr(keep[1])
    d.dir_y, d.dir_u = _np_ptr(fm.dir_y), _np_ptr(fm.dir_u)
    d.phi_y, d.grad_y, d.phi_u, d.wdet = _np_ptr(fm.phi_y), _np_ptr(fm.grad_y), _np_ptr(fm.phi_u), _np_ptr(fm.wdet)
    if with_coef and fm.coef.shape[1] == fm.ncells:
        d.ncoef, d.coef = 2, _np_ptr(fm.coef)
    else:
        d.ncoef, d.coef = 0, None
    for i in range(8):
        d.params[i] = fm.params[i]
    d.nfields_y, d.nfields_u = fm.nfy, fm.nfu
    keep += [np.ascontiguousarray(fm.field_nfree, np.int64), np.ascontiguousarray(fm.field_ndir, np.int64)]
    d.field_nfree, d.field_ndir = _np_ptr(keep[-2]), _np_ptr(keep[-1])
    if getattr(fm, "cell_jinvT", None) is not None:  # per-cell geometry (ABI v3)
        keep += [np.ascontiguousarray(fm.cell_jinvT, np.float64), np.ascontiguousarray(fm.cell_detj, np.float64)]
        d.cell_jinvT, d.cell_detj, d.geom_nq = _np_ptr(keep[-2]), _np_ptr(keep[-1]), int(fm.geom_nq)
    else:
        d.cell_jinvT, d.cell_detj, d.geom_nq = None, None, 0
    return d, keep


def dump_desc(fm, path: str, with_coef: bool = True):
    """Write a FlatModel as the flat binary file tests/c/test_sharded.c reads (a C program has no Gridap/Python host
    layer to flatten a model): magic, 10 int32, 5 int64, params[8], then every array of pdenlp_desc in declaration
    order (lengths follow from the scalars)."""
    d, keep = build_desc(fm, -1, MEM_HOST, with_coef)
    nfy, nfu = max(int(d.nfields_y), 1), (max(int(d.nfields_u), 1) if d.nu else 0)
    with open(path, "wb") as fh:
        fh.write(b"PDENLPD3")
        np.array([d.integrand, d.dim, d.ny, d.nu, d.nq, d.nparam, d.nfields_y, d.nfields_u, d.ncoef, d.geom_nq], np.int32).tofile(fh)
        np.array([d.ncells, d.nfree_y, d.nfree_u, d.ndir_y, d.ndir_u], np.int64).tofile(fh)
        np.array([d.params[i] for i in range(8)], np.float64).tofile(fh)
        arrays = [(keep[0], np.int32, d.ncells * nfy * d.ny), (keep[1], np.int32, d.ncells * nfu * d.nu),
                  (fm.dir_y, np.float64, d.ndir_y), (fm.dir_u, np.float64, d.ndir_u),
                  (fm.phi_y, np.float64, d.nq * d.ny), (fm.grad_y, np.float64, d.nq * d.ny * d.dim),
                  (fm.phi_u, np.float64, d.nq * d.nu), (fm.wdet, np.float64, d.nq),
                  (fm.coef if d.ncoef else np.zeros(0), np.float64, d.ncoef * d.ncells * d.nq),
                  (keep[9], np.int64, nfy + nfu), (keep[10], np.int64, nfy + nfu)]
        if d.geom_nq:
            arrays += [(keep[11], np.float64, d.ncells * d.geom_nq * d.dim * d.dim), (keep[12], np.float64, d.ncells * d.geom_nq)]
        for a, dt, n in arrays:
            a = np.ascontiguousarray(a, dt).reshape(-1)
            assert a.size == n, (a.size, n)
            a.tofile(fh)


class ShardedHandle:
    """Owns one ``pdenlp_sharded*``: the in-library multi-GPU handle (one process, host vectors in and out)."""

    def __init__(self, fm, ndev: int, devices=None, with_coef: bool = True):
        lib = load()
        d, keep = build_desc(fm, -1, MEM_HOST, with_coef)
        devs = (C.c_int32 * ndev)(*devices) if devices is not None else None
        h = C.c_void_p()
        check(lib.pdenlp_sharded_create(C.byref(d), ndev, devs, C.byref(h)))
        del keep
        self._h, self.ndev = h, ndev
        self.info = Info()
        check(lib.pdenlp_sharded_get_info(h, C.byref(self.info)))

    def close(self):
        if getattr(self, "_h", None):
            load().pdenlp_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def slab(self, k):
        a, b = C.c_int64(), C.c_int64()
        check(load().pdenlp_sharded_slab(self._h, k, C.byref(a), C.byref(b)))
        return a.value, b.value

    def obj(self, x) -> float:
        f = C.c_double()
        check(load().pdenlp_sharded_obj(self._h, _np_ptr(x), C.byref(f)))
        return f.value

    def _vec(self, fn, n, *args):
        out = np.empty(n)
        check(fn(self._h, *args, _np_ptr(out)))
        return out

    def grad(self, x):
        return self._vec(load().pdenlp_sharded_grad, self.info.nvar, _np_ptr(x))

    def cons(self, x):
        return self._vec(load().pdenlp_sharded_cons, self.info.ncon, _np_ptr(x))

    def jac_coord(self, x):
        return self._vec(load().pdenlp_sharded_jac_coord, self.info.nnzj, _np_ptr(x))

    def hess_coord(self, x, lam=None, obj_weight=1.0):
        return self._vec(load().pdenlp_sharded_hess_coord, self.info.nnzh, _np_ptr(x), _np_ptr(lam) if lam is not None else None, float(obj_weight))

    def jprod(self, x, v):
        return self._vec(load().pdenlp_sharded_jprod, self.info.ncon, _np_ptr(x), _np_ptr(v))

    def jtprod(self, x, v):
        return self._vec(load().pdenlp_sharded_jtprod, self.info.nvar, _np_ptr(x), _np_ptr(v))

    def hprod(self, x, lam, v, obj_weight=1.0):
        return self._vec(load().pdenlp_sharded_hprod, self.info.nvar, _np_ptr(x), _np_ptr(lam) if lam is not None else None, float(obj_weight), _np_ptr(v))

    def structure(self, hess: bool):
        n = self.info.nnzh if hess else self.info.nnzj
        r, c = np.empty(n, np.int64), np.empty(n, np.int64)
        fn = load().pdenlp_sharded_hess_structure if hess else load().pdenlp_sharded_jac_structure
        check(fn(self._h, _np_ptr(r), _np_ptr(c)))
        return r, c


class Handle:
    """Owns one ``pdenlp_model*`` created from a FlatModel (one cell slab on one GPU)."""

    def __init__(self, fm, device: int = -1, memspace: int = MEM_DEVICE, with_coef: bool = True, flags: int = 0):
        lib = load()
        d, keep = build_desc(fm, device, memspace, with_coef, flags)
        h = C.c_void_p()
        check(lib.pdenlp_create(C.byref(d), C.byref(h)))
        del keep
        self._h = h
        self.memspace = memspace
        info = Info()
        check(lib.pdenlp_get_info(h, C.byref(info)))
        self.info = info

    def close(self):
        if getattr(self, "_h", None):
            load().pdenlp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # thin pass-throughs; pointers are ints (device or host addresses)
    def set_stream(self, stream_ptr: int):
        check(load().pdenlp_set_stream(self._h, C.c_void_p(stream_ptr)))

    def sync(self):
        check(load().pdenlp_sync(self._h))

    def obj(self, x: int) -> float:
        f = C.c_double()
        check(load().pdenlp_obj(self._h, x, C.byref(f)))
        return f.value

    def grad(self, x, g):
        check(load().pdenlp_grad(self._h, x, g))

    def objgrad(self, x, g) -> float:
        f = C.c_double()
        check(load().pdenlp_objgrad(self._h, x, C.byref(f), g))
        return f.value

    def objcons(self, x, c) -> float:
        f = C.c_double()
        check(load().pdenlp_objcons(self._h, x, C.byref(f), c))
        return f.value

    def cons(self, x, c):
        check(load().pdenlp_cons(self._h, x, c))

    def jac_coord(self, x, vals):
        check(load().pdenlp_jac_coord(self._h, x, vals))

    def hess_coord(self, x, lam, obj_weight, vals):
        check(load().pdenlp_hess_coord(self._h, x, lam, float(obj_weight), vals))

    def hess_coord_obj(self, x, obj_weight, vals):
        check(load().pdenlp_hess_coord_obj(self._h, x, float(obj_weight), vals))

    def jprod(self, x, v, out):
        check(load().pdenlp_jprod(self._h, x, v, out))

    def jtprod(self, x, v, out):
        check(load().pdenlp_jtprod(self._h, x, v, out))

    def hprod(self, x, lam, obj_weight, v, out):
        check(load().pdenlp_hprod(self._h, x, lam, float(obj_weight), v, out))

    def jac_structure(self, rows, cols):
        check(load().pdenlp_jac_structure(self._h, rows, cols))

    def hess_structure(self, rows, cols):
        check(load().pdenlp_hess_structure(self._h, rows, cols))

    def input_ranges(self):
        """(y_lo, y_hi, u_lo, u_hi): the free-dof ranges the handle's cells touch (what a host-memspace model stages)."""
        r = (C.c_int64 * 4)()
        check(load().pdenlp_input_ranges(self._h, r))
        return tuple(int(v) for v in r)

    def launch_count(self) -> int:
        n = C.c_int64()
        check(load().pdenlp_launch_count(self._h, C.byref(n)))
        return n.value

    def template_stats(self) -> dict:
        st = (C.c_int64 * 4)()
        check(load().pdenlp_template_stats(self._h, st))
        return {"chunks": st[0], "template_cells": st[1], "generic_cells": st[2], "image_builds": st[3]}

    def graph_replays(self) -> int:
        n = C.c_int64()
        check(load().pdenlp_graph_replays(self._h, C.byref(n)))
        return n.value

    def selfcheck_error(self) -> float:
        e = C.c_double()
        check(load().pdenlp_selfcheck_error(self._h, C.byref(e)))
        return e.value

    def set_timing(self, on: bool):
        check(load().pdenlp_set_timing(self._h, 1 if on else 0))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(load().pdenlp_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class Compaction:
    """Owns one ``pdenlp_compaction*``: COO -> CSC/CSR plan with duplicates summed (device)."""

    def __init__(self, rows, cols, nrows: int, ncols: int, order: int = ORDER_CSC):
        """rows/cols: 1-based int64, numpy arrays (host) or CUDA torch tensors (device)."""
        lib = load()
        on_dev = int(hasattr(rows, "data_ptr"))
        pr = rows.data_ptr() if on_dev else rows.ctypes.data
        pc = cols.data_ptr() if on_dev else cols.ctypes.data
        n = int(rows.numel() if on_dev else rows.size)
        h = C.c_void_p()
        check(lib.pdenlp_compaction_create(n, pr, pc, nrows, ncols, order, on_dev, C.byref(h)))
        self._h = h
        self.nrows, self.ncols, self.order, self.nnz_in = nrows, ncols, order, n
        m = C.c_int64()
        check(lib.pdenlp_compaction_nnz(h, C.byref(m)))
        self.nnz = m.value

    def structure(self, rows_ptr=None, cols_ptr=None, major_ptr=None, on_device: bool = False):
        check(load().pdenlp_compaction_structure(self._h, rows_ptr, cols_ptr, major_ptr, int(on_device)))

    def apply(self, coo_vals_ptr: int, out_ptr: int, stream_ptr: int = 0):
        check(load().pdenlp_compaction_apply(self._h, coo_vals_ptr, out_ptr, C.c_void_p(stream_ptr)))

    def close(self):
        if getattr(self, "_h", None):
            load().pdenlp_compaction_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
