"""The C-ABI library loads on a CPU-only box, exports every symbol include/cimhead.h declares, and
validates its arguments before touching CUDA (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from cim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cimhead.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cim_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run `make -C cim_b200/csrc`"


def test_every_declared_symbol_is_exported_and_bound():
    handle = C.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in cimhead.h but not exported"
        assert n in _lib.EXPORTED_SYMBOLS, f"{n} has no ctypes signature in cim_b200/_lib.py"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_version_struct_layout_and_error_strings():
    L = _lib.lib()
    assert L.cim_abi_version() == _lib.ABI_VERSION
    assert L.cim_sizeof_mine_params() == C.sizeof(_lib.MineParams)
    assert L.cim_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert b"cimhead" in L.cim_error_string(code)


def test_argument_validation_needs_no_gpu():
    L = _lib.lib()
    null = C.c_void_p(0)
    assert L.cim_roi_align_fwd(null, null, null, 1, 1, 1, 1, 1, 7, 7, 1.0, 0, 1, null, 0, null) == -1
    assert L.cim_roi_pool_fwd(null, null, null, null, 1, 1, 1, 1, 1, 7, 7, 1.0, null) == -1
    assert L.cim_mask_pack(null, null, 1, 1, 1, null) == -1
    assert L.cim_mask_overlap(null, 1, 1, 1, null, null, null, null, null, 0, null) == -1
    assert L.cim_mask_overlap_workspace_bytes(8, 2000, 8192, 0) > 2 * 8 * 2000 * 2000 * 2
    assert L.cim_score_heads(null, null, null, null, 1, 1, 1, 1, 1, null, 0, null) == -1
    p = _lib.MineParams()
    assert L.cim_mine_workspace_bytes(C.byref(p)) == 0          # all-zero params are invalid
    p.n_img, p.R, p.C, p.C1, p.n_layers, p.det_cols, p.gt_cap, p.keep_count = 1, 300, 20, 21, 3, 21, 300, 30
    assert L.cim_mine_workspace_bytes(C.byref(p)) > 0
    p.R = 20000                                                  # beyond the documented limit
    assert L.cim_assign(C.byref(p), null, null, null, null, null, null, null, null, null, null, null) == -2
    assert L.cim_roi_align_workspace_bytes(16000) >= 16000 * 640


def test_no_cpu_fallback():
    import torch
    from cim_b200 import ops, mask_ops, heads
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.roi_align(torch.zeros(1, 32, 8, 8), torch.zeros(1, 5), 7)
    with pytest.raises(RuntimeError, match="no CPU path"):
        mask_ops.mask_pack(torch.zeros(2, 8, 8, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        heads.cls_iou_model(16, 21, 3)(torch.zeros(4, 16))


def test_debug_flag_constants_match_the_header():
    """cim_b200/_lib.py mirrors the CIM_DBG_* enum of include/cimhead.h by value; the flags round-trip through the
    library (cim_set_debug_flags / cim_get_debug_flags need no GPU)."""
    src = open(os.path.join(ROOT, "include", "cimhead.h")).read()
    header = {n: int(v) for n, v in re.findall(r"\bCIM_(DBG_[A-Z_]+)\s*=\s*(\d+)u", src)}
    assert len(header) >= 6
    for name, value in header.items():
        assert getattr(_lib, name) == value, f"_lib.{name} != CIM_{name} of cimhead.h"
    lib = _lib.lib()
    prev = lib.cim_get_debug_flags()
    try:
        with _lib.debug_flags(_lib.DBG_ROI_POOL_SIMPLE | _lib.DBG_SCORE_FFMA):
            assert lib.cim_get_debug_flags() == 36
        assert lib.cim_get_debug_flags() == prev
    finally:
        lib.cim_set_debug_flags(prev)
