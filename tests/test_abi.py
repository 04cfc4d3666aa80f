"""The C-ABI library loads on a CPU-only box and exports every symbol include/lgm_b200.h declares; argument errors
are reported through return codes + lgm_last_error_string (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lgm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgm_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge  # builds the library (nvcc cross-compiles without a GPU)
    ge.build()
    from lgm_b200 import _lib
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    from lgm_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lgm_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes signature table out of sync with the header"


def lib_step_counts_bytes():
    """sizeof(lgm_step_counts) as declared in the header: uint64 + 2 x uint32."""
    src = open(os.path.join(ROOT, "include", "lgm_b200.h")).read()
    body = re.search(r"typedef struct lgm_step_counts \{(.*?)\}", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    sizes = {"uint64_t": 8, "uint32_t": 4}
    return sum(sizes[t] for t in re.findall(r"(uint64_t|uint32_t)\s+\w+;", body))


def test_struct_layout_matches_header():
    from lgm_b200 import _lib
    assert ctypes.sizeof(_lib.RenderParams) == 8 * 4  # 5 int32 + 3 float
    assert lib_step_counts_bytes() == 16
    assert _lib.GRAD_ROW == 12


def test_size_queries_and_errors(lib):
    from lgm_b200 import _lib
    assert lib.lgm_abi_version() == 4
    assert lib.lgm_tiles_per_view(320, 320) == 400 and lib.lgm_tiles_per_view(512, 512) == 1024
    assert lib.lgm_tiles_per_view(17, 33) == 2 * 3
    assert lib.lgm_num_block_sums(98304, 208) == 208 * 384 and lib.lgm_num_block_sums(257, 3) == 6
    b = ctypes.c_size_t(0)
    ok = _lib.make_params(8, 98304, 208, 320, 320, 0.577, 0.577, 1.0)
    assert lib.lgm_bin_workspace_bytes(ok, 1_000_000, 0, b) == 0 and b.value >= 12_000_000
    assert lib.lgm_bin_workspace_bytes(ok, 1 << 30, 0, b) == -4  # LGM_ERR_TOO_MANY_INSTANCES
    assert b"2^30" in lib.lgm_last_error_string()
    bad = _lib.make_params(8, 98304, 208, 0, 320, 0.577, 0.577, 1.0)
    assert lib.lgm_bin_workspace_bytes(bad, 10, 0, b) == -2      # LGM_ERR_BAD_SHAPE
    bad = _lib.make_params(8, 98304, 208, 320, 320, 0.0, 0.577, 1.0)
    assert lib.lgm_bin_workspace_bytes(bad, 10, 0, b) == -5      # LGM_ERR_BAD_VALUE
    assert lib.lgm_bin_workspace_bytes(None, 10, 0, b) == -1     # LGM_ERR_NULL_POINTER
    huge = _lib.make_params(1, 1 << 30, 8, 320, 320, 0.5, 0.5, 1.0)
    assert lib.lgm_bin_workspace_bytes(huge, 10, 0, b) == -2
    assert lib.lgm_sort_workspace_bytes(4096 * 3 + 1, 49, b) == 0 and b.value >= 7 * 4 * 256 * 4
    assert lib.lgm_sort_workspace_bytes(10, 65, b) == -5
    assert lib.lgm_sort_input_is_tmp(48) == 0 and lib.lgm_sort_input_is_tmp(49) == 1
    # null pointers on entry points are rejected before any CUDA call
    assert lib.lgm_mark_visible(None, 5, None, None, None) == -1
    assert lib.lgm_mark_visible(None, -1, None, None, None) == -2
    assert lib.lgm_mark_visible(None, 0, None, None, None) == 0
    assert lib.lgm_sort_pairs(None, None, None, None, None, 5, 40, 0, None, 0) == -1
    assert lib.lgm_sort_pairs(None, None, None, None, None, 5, 64, 1, None, 0) == -5
    assert lib.lgm_forward_geom(None, ok, *([None] * 12)) == -1
    # the enqueue-only binning interface: sizes, modes, tuning names
    assert lib.lgm_count_workspace_bytes(ok, b) == 0 and b.value >= 2 * 4 * 208 * 400
    assert lib.lgm_direct_bin_tile_cap() == 20480
    assert lib.lgm_forward_count(None, ok, None, None, None, None, 0, None) == -1
    assert lib.lgm_forward_bin(None, ok, None, None, None, None, 10, -1, 0, 7, None, None, None, None, 0, None, 0) == -5  # bin_mode
    too_many_views = _lib.make_params(1, 16, 65536, 32, 32, 0.5, 0.5, 1.0)
    assert lib.lgm_bin_workspace_bytes(too_many_views, 10, 0, b) == -2 and b"65535" in lib.lgm_last_error_string()
    assert lib.lgm_set_tuning(b"fwd_batch", 256) == 0 and lib.lgm_set_tuning(b"fwd_batch", -1) == 0
    assert lib.lgm_set_tuning(b"no_such_knob", 1) == -5
    assert lib.lgm_activate_forward(None, 2, 100, None, None, 0, None) == -1   # reference axis needs its scratch
    assert lib.lgm_activate_forward(None, 2, 100, None, None, 5, None) == -5   # rot_axis
    assert lib.lgm_backward_composite(None, ok, *([None] * 14)) == -1
