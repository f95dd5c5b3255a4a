"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/biolith_b200.h declares, and the host wrapper fails loudly (no fallback) without a GPU."""

import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "biolith_b200.h")).read()
    return sorted(set(re.findall(r"BL_API\s+[\w\s\*]+?\b(bl_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from biolith_b200 import _lib

    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"
    assert lib.bl_version() == _lib.BL_ABI_VERSION
    assert lib.bl_strerror(-2).decode().startswith("option outside")


def test_struct_layout_matches_header():
    import ctypes as C

    from biolith_b200._lib import bl_desc, bl_info

    # int32 x6, int64, int32 x8, double x7  (natural alignment, no packing pragma)
    assert C.sizeof(bl_desc) == 6 * 4 + 8 + 8 * 4 + 7 * 8
    assert C.sizeof(bl_info) == 2 * 4 + 4 * 8 + 2 * 4


def test_no_cpu_fallback_without_gpu():
    import biolith_b200
    from biolith_b200 import _lib

    lib = _lib.load()
    import ctypes as C

    n = C.c_int(0)
    rc = lib.bl_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    with pytest.raises(biolith_b200.BiolithB200Error):
        biolith_b200.OccupancyLikelihood("occu", np.zeros((4, 1)), np.zeros((4, 1, 3, 1)), np.zeros((1, 4, 1, 3)))


def test_product_does_not_import_oracle_or_torch():
    pkg = os.path.join(ROOT, "biolith_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert not re.search(r"^\s*(from|import)\s+(torch|triton)\b", src, re.M), f"{f} imports torch/triton"


def test_argument_validation_needs_no_device():
    """NULL handles / outputs are rejected before any CUDA call (same status on CPU-only hosts and GPU boxes)."""
    import ctypes as C

    from biolith_b200 import _lib

    lib = _lib.load()
    k = C.c_int32(-1)
    assert lib.bl_plan_kernel(None, 5, C.byref(k), None, None, None) == -1   # BL_ERR_INVALID
    assert lib.bl_last_error().decode() != ""
    ms = C.c_float()
    assert lib.bl_eval_timed(None, None, 1, None, None, None, 1, C.byref(ms)) != 0
    assert lib.bl_dataset_info(None, None) != 0
