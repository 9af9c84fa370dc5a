"""The C-ABI library loads and exports every symbol include/bmo.h declares (no compute without a GPU),
and fails loudly -- never falls back to a CPU path -- when no sm_100 device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "bmo.h")).read()
    return sorted(set(re.findall(r"\b(bmo_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(bmo):
    from bmo_b200 import _lib
    lib = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bmo.h but not exported by libbmo.so"
    assert set(_lib.EXPORTS) == set(syms)


def test_struct_layouts_match_header(bmo):
    from bmo_b200 import _lib
    assert C.sizeof(_lib.bmo_prim) == 8 + 8 * (3 + 9 + 4) + 8
    assert C.sizeof(_lib.bmo_part) == 6 * 4 + 8 * (2 + 10)
    assert C.sizeof(_lib.bmo_object) == 4 * 4 + 8 * (3 + 9 + 2)
    assert C.sizeof(_lib.bmo_mesh) == 4 * 8 + 8


def test_no_cpu_fallback(bmo):
    """Without a GPU the product path must raise (the oracle is never used as a fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tests import scenes
    sc = scenes.doublet_spot(bmo)
    with pytest.raises(bmo.BmoError):
        bmo.solve_system_(sc["system"], bmo.Beam((0, -0.05, 0), (0, 1, 0), 707e-9))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "beamletoptics.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("test infrastructure", ""), f"{f} mentions the oracle"
