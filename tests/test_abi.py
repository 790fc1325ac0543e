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


def test_glv_decomposition_host_arithmetic():
    """csrc/glv.h through g16_glv_decompose (pure host arithmetic): k = k1 + k2 * lambda (mod r) with both magnitudes
    below 2^127, for edge scalars and random ones; lambda and beta are a matching pair of cube roots of unity
    (phi(G) = (beta x, y) = lambda G in the oracle's group law); scalars >= 2^254 that do not split are refused."""
    import ctypes as C
    import random
    import g16_oracle as o
    from g16b200 import _lib
    lib = _lib.load()
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    beta = 0x59e26bcea0d48bacd4f263f1acdb5c4f5763473177fffffe
    assert pow(lam, 3, o.R) == 1 and pow(beta, 3, o.P) == 1 and lam != 1 and beta != 1
    assert o.g1_mul(lam, o.GEN1) == (beta * o.GEN1[0] % o.P, o.GEN1[1])

    def dec(k):
        kk = (C.c_uint64 * 4)(*[(k >> (64 * i)) & (2 ** 64 - 1) for i in range(4)])
        a, b, n1, n2 = (C.c_uint64 * 2)(), (C.c_uint64 * 2)(), C.c_int(), C.c_int()
        rc = lib.g16_glv_decompose(kk, a, b, C.byref(n1), C.byref(n2))
        k1, k2 = a[0] | (a[1] << 64), b[0] | (b[1] << 64)
        return rc, (-k1 if n1.value else k1), (-k2 if n2.value else k2)

    rnd = random.Random(11)
    edge = [0, 1, 2, o.R - 1, o.R - 2, lam, lam - 1, lam + 1, 1 << 253, o.R // 2, o.R // 3, (o.R - 1) // 2,
            (1 << 127) - 1, 1 << 127, 1 << 128]
    for t in range(20000):
        k = edge[t] if t < len(edge) else rnd.randrange(o.R)
        rc, k1, k2 = dec(k)
        assert rc == 0 and (k1 + k2 * lam - k) % o.R == 0, hex(k)
        assert abs(k1) < 1 << 127 and abs(k2) < 1 << 127
    # the split also drives a scalar multiplication correctly (what the device does with its doubling table)
    for k in (edge[3], rnd.randrange(o.R), rnd.randrange(o.R)):
        _, k1, k2 = dec(k)
        p = o.g1_mul(rnd.randrange(1, o.R), o.GEN1)
        phi = (beta * p[0] % o.P, p[1])
        a = o.g1_mul(abs(k1), p if k1 >= 0 else o.g1_neg(p))
        b = o.g1_mul(abs(k2), phi if k2 >= 0 else o.g1_neg(phi))
        assert o.g1_add(a, b) == o.g1_mul(k, p)
