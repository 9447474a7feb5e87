"""Shared helpers for the parity tests."""
import numpy as np


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_batches_identical(got, ref, what=""):
    """Discrete outputs and coordinates bit-identical (NaNs compare by bit pattern class)."""
    assert np.array_equal(got.iters, ref.iters), f"{what}: iters differ at {np.nonzero(got.iters != ref.iters)}"
    assert np.array_equal(got.converged, ref.converged), f"{what}: converged flags differ"
    assert np.array_equal(got.root_index, ref.root_index), f"{what}: root index differs"
    for c, (a, b) in enumerate(zip(got.out, ref.out)):
        same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"{what}: out[{c}] differs at {np.nonzero(~same)[0][:8]}"
    if got.cand is not None and ref.cand is not None:
        same = (bits(got.cand) == bits(ref.cand)) | (np.isnan(got.cand) & np.isnan(ref.cand))
        assert same.all(), f"{what}: candidates differ"


def rel_err(a, b, scale):
    """|a-b| / max(1, |b|, scale): SURVEY section 7's relative-error definition."""
    den = np.maximum(np.maximum(1.0, np.abs(b)), scale)
    return np.abs(a - b) / den
