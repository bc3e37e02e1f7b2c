"""The parity bar of the GPU tests, in one place (BASELINE.json north_star: "values must match within a stated
relative tolerance of 1e-12 in FP64").

Two bounds per compared array (COO segment or dof-indexed vector), both RTOL = 1e-12:
  * element-wise: |got - ref| <= RTOL * |ref| for every entry with |ref| > ELEM_FLOOR * max|ref|;
  * segment-wise: |got - ref| <= RTOL * max|ref| for the rest — entries far below the segment's scale are sums
    with cancellation (or exact zeros in one evaluation order and round-off in the other) and cannot meet an
    element-wise bound in any arithmetic.
"""
import numpy as np

RTOL = 1e-12
ELEM_FLOOR = 1e-3


def parity_metrics(got, ref):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape:
        return float("inf"), float("inf")
    if ref.size == 0:
        return 0.0, 0.0
    scale = float(np.abs(ref).max())
    if scale == 0.0:
        return float(np.abs(got).max()), 0.0
    d = np.abs(got - ref)
    big = np.abs(ref) > ELEM_FLOOR * scale
    return float(d.max() / scale), (float((d[big] / np.abs(ref[big])).max()) if big.any() else 0.0)


def is_close(got, ref, rtol=RTOL):
    seg, elem = parity_metrics(got, ref)
    return seg <= rtol and elem <= rtol


def assert_parity(got, ref, what="", rtol=RTOL):
    seg, elem = parity_metrics(got, ref)
    assert seg <= rtol and elem <= rtol, f"{what}: segment-relative error {seg:.3e}, element-wise relative error {elem:.3e} (bar {rtol:g})"
