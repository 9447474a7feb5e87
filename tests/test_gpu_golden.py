"""CUDA path against the golden vectors generated from the reference's own sources."""
import numpy as np
import pytest

from test_golden import check_against_golden, load_numeric

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", [1, 2, 3, 4, 5])
def test_cuda_matches_reference_golden(gpu, gcs, kind, variant):
    hb, z = load_numeric(gcs.capi, kind)
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    check_against_golden(hb, z, kind, f"cuda kind {kind} variant {variant}")
