"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gcs_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gcs_b200.h")).read()
    return sorted(set(re.findall(r"GCS_B200_API\s+[\w\s\*]+?\b(gcs_b200_\w+)\s*\(", text)))


def test_header_symbols_are_all_exported(gcs, built):
    lib = gcs.capi.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in gcs_b200.h but not exported"
    assert set(names) == set(gcs.capi.EXPORTS)


def test_kind_tables_match_python_mirror(gcs, built):
    lib, capi = gcs.capi.load(), gcs.capi
    for k in range(0, 8):
        assert lib.gcs_b200_kind_in_cols(k) == capi.IN_COLS.get(k, 0)
        assert lib.gcs_b200_kind_out_cols(k) == capi.OUT_COLS.get(k, 0)
    assert b"sm_100a" in lib.gcs_b200_version()


def test_batch_struct_layout_matches_header(gcs):
    capi = gcs.capi
    # int32 x2, int64, int32 x2, 13 ptr, ptr, ptr, 4 ptr, 4 ptr
    assert C.sizeof(capi.CBatch) == 4 + 4 + 8 + 4 + 4 + 8 * (13 + 1 + 1 + 4 + 4)
    assert capi.CBatch.in_.offset == 24 and capi.CBatch.code.offset == 24 + 13 * 8


def test_no_cpu_fallback_without_a_device(gcs, built):
    """In the CPU container the compute entry points must refuse, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    capi = gcs.capi
    lib = capi.load()
    assert lib.gcs_b200_device_count() == 0
    hb = gcs.synth.make_pp(8).alloc_outputs()
    cb = hb.cbatch()
    assert lib.gcs_b200_solve_host(C.byref(cb), 0) == capi.GCS_E_NO_DEVICE
    assert b"no CPU fallback" in lib.gcs_b200_last_error()
    assert np.isnan(hb.out[0]).all()  # nothing was written
    with pytest.raises(capi.GcsError):
        capi.solve_host(hb, 0)
    assert lib.gcs_b200_fp64_probe(0, 0) < 0
    lib.gcs_b200_host_alloc.restype = C.c_void_p
    lib.gcs_b200_host_alloc.argtypes = [C.c_size_t]
    assert lib.gcs_b200_host_alloc(4096) is None  # no device: the caller falls back to ordinary memory


def test_null_columns_only_where_the_shape_is_anchored(gcs, built):
    """A NULL input column is a column of zeros, allowed only for the anchor columns of the
    zero-fixed solvers (gcs_b200.h); anything else is GCS_E_INVALID - checked before any device is
    touched, so it runs here."""
    capi, synth = gcs.capi, gcs.synth
    lib = capi.load()
    allowed = {1: [0, 1, 4], 2: [0, 1, 3], 3: [], 4: [], 5: [1, 7, 10, 11]}
    for kind, cols in allowed.items():
        assert [c for c in range(capi.IN_COLS[kind]) if lib.gcs_b200_column_may_be_null(kind, c)] == cols
        for c in range(capi.IN_COLS[kind]):
            hb = synth.make(kind, 4).alloc_outputs()
            hb.cols[c] = None
            cb = hb.cbatch()
            rc = lib.gcs_b200_solve_host(C.byref(cb), 0)
            if c in cols:
                assert rc != capi.GCS_E_INVALID  # accepted by validation (then: no device here, or solved)
            else:
                assert rc == capi.GCS_E_INVALID and b"anchor" in lib.gcs_b200_last_error()


def test_argument_checks_of_the_round_2_entry_points_need_no_device(gcs, built):
    """Validation comes before any device is touched: bad index ranges, bad batch lists, host pointers
    handed to the device entry points and unknown variants are GCS_E_INVALID here as on a GPU box."""
    capi, synth = gcs.capi, gcs.synth
    lib = capi.load()
    hb = synth.make_pp(16).alloc_outputs()
    cb = hb.cbatch()
    for first, count in ((-1, 4), (0, 17), (12, 5), (17, 0), (3, -1)):
        assert lib.gcs_b200_solve_host_range_async(C.byref(cb), 0, first, count) == capi.GCS_E_INVALID, (first, count)
        assert b"index range" in lib.gcs_b200_last_error()
    assert lib.gcs_b200_solve_many(None, 0, 0, None) == capi.GCS_OK          # an empty job is a no-op anywhere
    assert lib.gcs_b200_solve_many(None, 2, 0, None) == capi.GCS_E_INVALID
    arr = (C.POINTER(capi.CBatch) * 2)(C.pointer(cb), C.pointer(cb))
    assert lib.gcs_b200_solve_many(arr, 2, 0, None) == capi.GCS_E_INVALID    # host pointers
    assert b"device pointers" in lib.gcs_b200_last_error()
    cb.variant = 11
    assert lib.gcs_b200_solve_host(C.byref(cb), 0) == capi.GCS_E_INVALID and b"unknown variant" in lib.gcs_b200_last_error()
    # kernel mapping a class resolves to: the bit-identical K4 and the 8-seed K1 take the sequential kernel, the contracted K4 the linear one
    R = lib.gcs_b200_resolve_variant
    assert R(capi.VARIANT_DEFAULT, 1, 1 << 19, 2) == capi.VARIANT_SORTED and R(capi.VARIANT_DEFAULT, 1, 1000, 2) == capi.VARIANT_STATIC
    assert R(capi.VARIANT_DEFAULT, 4, 1 << 19, 2) == capi.VARIANT_SEQ and R(capi.VARIANT_CONTRACTED, 4, 1 << 19, 2) == capi.VARIANT_CONTRACTED_LINEAR
    assert R(capi.VARIANT_CONTRACTED, 1, 1 << 19, 2) == capi.VARIANT_CONTRACTED_STATIC
    assert R(capi.VARIANT_CONTRACTED, 1, 1 << 19, 8) == capi.VARIANT_CONTRACTED_SEQ
    assert R(capi.VARIANT_CONTRACTED, 4, 100, 8) == capi.VARIANT_CONTRACTED_LINEAR and R(capi.VARIANT_CONTRACTED_LINEAR, 1, 1 << 19, 2) == capi.VARIANT_CONTRACTED_STATIC
    assert R(capi.VARIANT_REFILL, 3, 5, 2) == capi.VARIANT_REFILL
    assert lib.gcs_b200_pcie_probe(0, 0, 1, 1, 0, 1, (C.c_double * 4)()) == capi.GCS_E_INVALID


def test_library_is_built_from_the_sources_in_the_tree(gcs, built):
    """gcs_b200_version() carries the hash of csrc/*, the header and the flags it was compiled
    from: a stale libgcs_b200.so (the .so is not in git but travels to the GPU box) fails here."""
    import __graft_entry__ as g
    assert g.cuda_source_hash().encode() in gcs.capi.load().gcs_b200_version()


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may name it."""
    pkg = os.path.join(ROOT, "2d_geometry_constraint_solver_b200")
    bad = []
    for base in (pkg, os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c")) or f == "Makefile":
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"gcs_oracle|oracle_lib|libgcs_oracle|/oracle/|\.\./oracle", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
