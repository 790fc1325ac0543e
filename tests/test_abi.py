"""CPU: the C-ABI library builds, loads, exports every symbol include/g16b200.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "g16b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g16_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from g16b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing symbol " + n
        assert n in _lib.SIGNATURES, "binding missing for " + n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.g16_version() == 2


def test_struct_layouts_match_header():
    from g16b200 import _lib
    assert ctypes.sizeof(_lib.ProofRaw) == 256
    assert ctypes.sizeof(_lib.Stats) == 48
    assert ctypes.sizeof(_lib.Toxic) == 160
    assert ctypes.sizeof(_lib.ZkeyView) == 8 * 4 + 8 + 6 * 8 + (8 + 8 + 16 + 8 + 16) * 8
    assert _lib.PARTIALS_BYTES == 400


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import g16b200
    with pytest.raises(g16b200._lib.G16Error, match="no CPU fallback|CUDA"):
        g16b200.msm_g1(np.zeros((1, 4), np.uint64), np.zeros((1, 8), np.uint64))
    with pytest.raises(g16b200._lib.G16Error):
        g16b200.forward_ntt(np.zeros((2, 4), np.uint64), g16b200.create_domain(2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nim-groth16_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".nim")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "g16_oracle" not in txt and "oracle/" not in txt.replace("closed-form oracle", ""), f
