"""pdenlp_b200 — B200-native derivative-evaluation hot path for PDENLPModels.jl.

The directory is named ``pdenlpmodels.jl_b200`` (not an importable identifier);
``__graft_entry__.load_package()`` registers it as the module ``pdenlp_b200``.

Layout (only what the hot path needs):
  csrc/         CUDA kernels (sm_100a, FP64) + the C ABI -> libpdenlp_cuda.so
  capi.py       ctypes binding of include/pdenlp_cuda.h
  model.py      host-side mirror of the reference interface (GridapPDENLPModel,
                NLPModels callbacks, DimensionError, Counters)
  cartesian.py  stand-in for Gridap's host-side setup (mesh, spaces, numbering)
  flatten.py    flattening layer: FE objects -> pdenlp_desc arrays + COO structure
  problems.py   the closed set of named integrands / BASELINE configs
  bounds.py     bounds_functions_to_vectors (lvar / uvar from functions or vectors)
  sharded.py    cell-slab partition over one process per GPU
"""
from . import cartesian, problems, flatten, bounds  # noqa: F401
from .bounds import bounds_functions_to_vectors  # noqa: F401
from .problems import (  # noqa: F401
    controlelasticmembrane, poissonboltzman2d, burger1d, burger1d_param, poissonmixed, penalizedpoisson,
    basicunconstrained, controlsir, cellincrease, dynamicsir, torebrachistochrone, poissonparam, sis, poissonmixed2, BASELINE_CONFIGS,
)

__all__ = ["cartesian", "problems", "flatten", "bounds"]


def __getattr__(name):
    # the CUDA-facing modules are imported lazily so that the host-only pieces
    # (numbering, flattening) stay usable where libpdenlp_cuda.so is not built
    if name in ("capi", "model", "sharded", "build"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("GridapPDENLPModel", "DimensionError"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
