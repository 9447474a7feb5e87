"""CUDA path against the golden vectors generated from the reference's own sources."""
import numpy as np
import pytest

from test_golden import check_against_golden, check_against_golden_contract, load_numeric

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 9])
@pytest.mark.parametrize("kind", [1, 2, 3, 4, 5])
def test_cuda_matches_reference_golden(gpu, gcs, kind, variant):
    hb, z = load_numeric(gcs.capi, kind)
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    check_against_golden(hb, z, kind, f"cuda kind {kind} variant {variant}")


@pytest.mark.parametrize("variant", [5, 6, 7, 8])
@pytest.mark.parametrize("kind", [1, 2, 3, 4, 5])
def test_contracted_cuda_matches_reference_golden_to_the_contract(gpu, gcs, kind, variant):
    """The contracted variants against the golden vectors of the reference's own code: iteration
    counts, convergence flags and the chosen root equal; candidates and the chosen result within
    1e-9 relative (the north star's tolerance), NaN where the reference has NaN."""
    hb, z = load_numeric(gcs.capi, kind)
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    check_against_golden_contract(hb, z, kind, f"cuda kind {kind} variant {variant}")
