"""The checker of the north star's contract (util.assert_batches_within_contract, used for the
contracted kernel variants) must itself reject what the contract forbids: a different iteration
count, flag or root; a coordinate off by more than 1e-9 relative; a NaN where the reference has
none.  CPU only: the batches come from the oracle."""
import copy

import numpy as np
import pytest

import oracle_lib as O
from util import assert_batches_within_contract


@pytest.fixture(scope="module")
def pair(gcs, built):
    ref = O.solve(gcs.synth.make(1, 512).alloc_outputs())
    return ref


def _clone(b):
    c = copy.copy(b)
    c.out = [o.copy() for o in b.out]
    c.cand = b.cand.copy()
    c.iters, c.converged, c.root_index = b.iters.copy(), b.converged.copy(), b.root_index.copy()
    return c


def test_identical_and_last_bit_noise_pass(pair):
    got = _clone(pair)
    assert assert_batches_within_contract(got, pair) == 0.0
    got.out[0] = np.nextafter(got.out[0], np.inf)
    got.cand[0, 0] *= 1.0 + 1e-13
    assert assert_batches_within_contract(got, pair) < 1e-12


@pytest.mark.parametrize("field", ["iters", "converged", "root_index"])
def test_any_discrete_difference_fails(pair, field):
    got = _clone(pair)
    a = getattr(got, field)
    a.reshape(-1)[7] ^= 1
    with pytest.raises(AssertionError):
        assert_batches_within_contract(got, pair)


def test_coordinate_error_beyond_the_tolerance_fails(pair):
    got = _clone(pair)
    scale = max(1.0, float(np.max(np.abs(np.stack(pair.cols))[:, 3])), abs(float(pair.out[1][3])))
    got.out[1][3] += 3e-9 * scale
    with pytest.raises(AssertionError):
        assert_batches_within_contract(got, pair)
    got = _clone(pair)
    got.out[1][3] += 1e-10 * scale
    assert_batches_within_contract(got, pair)


def test_a_nan_the_reference_does_not_have_fails(pair):
    got = _clone(pair)
    got.cand[1, 0, 11] = np.nan
    with pytest.raises(AssertionError):
        assert_batches_within_contract(got, pair)
