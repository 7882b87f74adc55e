"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol
include/karios_b200.h declares, reports errors through kr_last_error, and the
ctypes structs match the header layout.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from karios_b200 import _native
    return _native.load_library()


def test_header_symbols_are_exported(lib):
    from karios_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "karios_b200.h")).read()
    declared = set(re.findall(r"KR_API\s+(?:const\s+)?\w+\s*\*?\s*(kr_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_error_channel(lib):
    assert lib.kr_version() >= 100
    out = C.c_void_p()
    rc = lib.kr_ctx_create(0, 0, 0, 0, C.byref(out))          # invalid size: fails before any CUDA call
    assert rc == -1 and not out.value
    assert b"out of range" in lib.kr_last_error()


def test_struct_layouts_match_header():
    from karios_b200 import _native as N
    assert C.sizeof(N.KltConf) == 12 * 4 + 6 * 8                # 11 int32 + padding
    assert N.KltConf.compute_mi.offset == 40 and N.KltConf.quality_level.offset == 48
    assert C.sizeof(N.Stats) == 4 * 8 + 8 + 4 + 14 * 4 + 4      # padded to 8
    assert N.Stats.est_cut_bits.offset == 92 and N.Stats.rows_skipped.offset == 96
    assert C.sizeof(N.UnitHeader) == 128 and N.UnitHeader.n.offset == 8 and N.UnitHeader.min_dx.offset == 48
    assert N.Stats.eig_max.offset == 40 and N.Stats.select_incomplete.offset == 72
    assert C.sizeof(N.Rows) == 8 * 8 + 8 and N.Rows.capacity.offset == 64


def test_no_cpu_fallback_without_gpu():
    import torch
    from karios_b200 import _native as N
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.KariosB200Error):
        N.Context(64, 64, 100)
    import numpy as np
    from karios_b200.matcher.klt import klt_tracker
    from karios_b200.core.configuration import KLTConfiguration
    z = np.zeros((32, 32), np.uint8)
    with pytest.raises(Exception):
        klt_tracker(z, z, None, KLTConfiguration())


def test_product_does_not_import_oracle():
    """The product package must never import / execute anything under oracle/."""
    pkg = os.path.join(ROOT, "karios_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert not re.search(r"#\s*include\s*[<\"].*oracle", src), f
                assert "libklt_oracle" not in src, f
