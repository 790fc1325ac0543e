"""CPU: the product's field / curve / digit headers compiled with g++ (portable path of field.cuh) against
the oracle.  This checks the formulas the CUDA kernels are built from; the PTX carry chains themselves are
checked against this same portable path on the GPU (g16_selftest, tests/test_gpu_core.py)."""
import ctypes
import os
import random
import subprocess

import pytest

import g16_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "hostemu.cpp")
LIB = os.path.join(HERE, "hostemu", "libhostemu.so")
CSRC = os.path.join(HERE, "..", "nim-groth16_b200", "csrc")


# the second build takes the whole-Fp2-call layer the G2 MSM is compiled with (field.cuh, G16_FP2_WHOLE_CALL)
@pytest.fixture(scope="module", params=["default", "fp2_whole_call"])
def L(request):
    lib = LIB if request.param == "default" else LIB.replace(".so", "_whole.so")
    flags = [] if request.param == "default" else ["-DG16_FP2_WHOLE_CALL"]
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", CSRC] + flags + ["-o", lib, SRC])
    return ctypes.CDLL(lib)


def buf(x):
    return (ctypes.c_uint8 * len(x)).from_buffer_copy(x)


def out(n=32):
    return (ctypes.c_uint8 * n)()


def le(x):
    return x.to_bytes(32, "little")


def test_field_ops(L):
    rnd = random.Random(1)
    rinv_r, rinv_p = pow(o.MONT, -1, o.R), pow(o.MONT, -1, o.P)
    for t in range(1500):
        a = rnd.randrange(o.R) if t > 2 else o.R - 1
        b = rnd.randrange(o.R) if t > 1 else o.R - 1
        c = out(); L.he_fr_mul(buf(le(a)), buf(le(b)), c)
        assert int.from_bytes(bytes(c), "little") == a * b * rinv_r % o.R
        a = rnd.randrange(o.P) if t > 2 else o.P - 1
        b = rnd.randrange(o.P) if t > 1 else o.P - 1
        c = out(); L.he_fp_mul(buf(le(a)), buf(le(b)), c)
        assert int.from_bytes(bytes(c), "little") == a * b * rinv_p % o.P
        c = out(); L.he_fp_add(buf(le(a)), buf(le(b)), c)
        assert int.from_bytes(bytes(c), "little") == (a + b) % o.P
        c = out(); L.he_fp_sub(buf(le(a)), buf(le(b)), c)
        assert int.from_bytes(bytes(c), "little") == (a - b) % o.P
    a = rnd.randrange(o.R)
    c = out(); L.he_fr_inv(buf(o.fr_to_mont_bytes(a)), c)
    assert o.fr_from_mont_bytes(bytes(c)) == pow(a, -1, o.R)
    c = out(); L.he_fr_halve(buf(le(a)), c)
    assert int.from_bytes(bytes(c), "little") == a * o.ONE_HALF_FR % o.R
    c = out(); L.he_fr_to_mont(buf(le(a)), c)
    assert bytes(c) == o.fr_to_mont_bytes(a)
    c = out(); L.he_fr_from_mont(buf(o.fr_to_mont_bytes(a)), c)
    assert int.from_bytes(bytes(c), "little") == a


def test_fp2_ops(L):
    rnd = random.Random(2)
    enc = lambda x: o.fp_to_mont_bytes(x[0]) + o.fp_to_mont_bytes(x[1])
    dec = lambda b: (o.fp_from_mont_bytes(b[:32]), o.fp_from_mont_bytes(b[32:]))
    for _ in range(100):
        x = (rnd.randrange(o.P), rnd.randrange(o.P))
        y = (rnd.randrange(o.P), rnd.randrange(o.P))
        c = out(64); L.he_fp2_mul(buf(enc(x)), buf(enc(y)), c)
        assert dec(bytes(c)) == o.fp2_mul(x, y)
        c = out(64); L.he_fp2_sqr(buf(enc(x)), c)
        assert dec(bytes(c)) == o.fp2_sqr(x)
    c = out(64); L.he_fp2_inv(buf(enc(x)), c)
    assert dec(bytes(c)) == o.fp2_inv(x)


def test_g1_group_law_with_special_cases(L):
    rnd = random.Random(3)
    pts = [o.g1_mul(rnd.randrange(1, o.R), o.GEN1) for _ in range(10)]
    pts += [pts[0], o.g1_neg(pts[1]), o.INF_G1, pts[2], pts[2]]          # duplicates, negations, infinity
    neg = [rnd.randrange(2) for _ in pts]
    exp = o.INF_G1
    for p, s in zip(pts, neg):
        exp = o.g1_add(exp, o.g1_neg(p) if s else p)
    pb = b"".join(o.g1_to_bytes(p) for p in pts)
    c = out(64); L.he_g1_sum(buf(pb), (ctypes.c_int * len(neg))(*neg), len(pts), c)
    assert o.g1_from_bytes(bytes(c)) == exp
    for seq, negs in (([pts[3], pts[3]], [0, 1]), ([pts[3], pts[3]], [0, 0]), ([o.INF_G1], [0])):
        pb2 = b"".join(o.g1_to_bytes(p) for p in seq)
        e = o.INF_G1
        for p, s in zip(seq, negs):
            e = o.g1_add(e, o.g1_neg(p) if s else p)
        c = out(64); L.he_g1_sum(buf(pb2), (ctypes.c_int * len(negs))(*negs), len(seq), c)
        assert o.g1_from_bytes(bytes(c)) == e
    c = out(64); L.he_g1_treesum(buf(pb), len(pts), c)
    e = o.INF_G1
    for p in pts:
        e = o.g1_add(e, p)
    assert o.g1_from_bytes(bytes(c)) == e
    k = rnd.randrange(o.R)
    c = out(64); L.he_g1_scalar_mul(buf(le(k)), buf(o.g1_to_bytes(pts[0])), c)
    assert o.g1_from_bytes(bytes(c)) == o.g1_mul(k, pts[0])
    c = out(64); L.he_g1_mul_u32(ctypes.c_uint32(54321), buf(o.g1_to_bytes(pts[0])), c)
    assert o.g1_from_bytes(bytes(c)) == o.g1_mul(54321, pts[0])


def test_g2_group_law(L):
    rnd = random.Random(4)
    q = [o.g2_mul(rnd.randrange(1, o.R), o.GEN2) for _ in range(4)]
    q += [q[0], o.g2_neg(q[1]), o.INF_G2, q[2], q[2]]
    neg = [rnd.randrange(2) for _ in q]
    exp = o.INF_G2
    for p, s in zip(q, neg):
        exp = o.g2_add(exp, o.g2_neg(p) if s else p)
    pb = b"".join(o.g2_to_bytes(p) for p in q)
    c = out(128); L.he_g2_sum(buf(pb), (ctypes.c_int * len(neg))(*neg), len(q), c)
    assert o.g2_from_bytes(bytes(c)) == exp
    c = out(128); L.he_g2_treesum(buf(pb), len(q), c)
    e = o.INF_G2
    for p in q:
        e = o.g2_add(e, p)
    assert o.g2_from_bytes(bytes(c)) == e
    k = rnd.randrange(o.R)
    c = out(128); L.he_g2_scalar_mul(buf(le(k)), buf(o.g2_to_bytes(q[0])), c)
    assert o.g2_from_bytes(bytes(c)) == o.g2_mul(k, q[0])


def test_signed_digits(L):
    rnd = random.Random(5)
    for c_ in (2, 4, 7, 13, 16, 17, 20, 22):
        nw = (255 + c_ - 1) // c_
        for t in range(100):
            k = rnd.randrange(o.R) if t > 3 else (o.R - 1, 0, 1, (1 << 254) - 1)[t]
            d = (ctypes.c_int * nw)(); L.he_digits(buf(le(k)), c_, nw, d)
            assert sum(int(d[w]) << (c_ * w) for w in range(nw)) == k
            assert all(-(1 << (c_ - 1)) < d[w] <= (1 << (c_ - 1)) for w in range(nw))


def test_ntt_pass_plan_emulation(L):
    """The multi-pass tile/butterfly index arithmetic of ntt_plan.cuh, emulated on the host, must
    reproduce forwardNTT / inverseNTT (ntt.nim) for every size the planner splits differently."""
    rnd = random.Random(6)
    for lg in (1, 2, 5, 11, 12, 13, 14):
        n = 1 << lg
        xs = [rnd.randrange(o.R) for _ in range(n)]
        D = o.create_domain(n)
        data = b"".join(o.fr_to_mont_bytes(x) for x in xs)
        for inverse in (0, 1):
            res = out(32 * n)
            L.he_ntt(buf(data), res, lg, inverse)
            got = [o.fr_from_mont_bytes(bytes(res)[32 * i:32 * i + 32]) for i in range(n)]
            want = o.inverse_ntt_fast(xs, D) if inverse else o.forward_ntt_fast(xs, D)
            assert got == want, (lg, inverse)


def test_shift_eval_domain_emulation(L):
    """prover.nim:109-113 shiftEvalDomain through the DIF -> coset[bitrev] -> DIT structure of ntt.cu."""
    rnd = random.Random(7)
    for lg in (1, 3, 11, 12, 13):
        n = 1 << lg
        xs = [rnd.randrange(o.R) for _ in range(n)]
        D = o.create_domain(n)
        eta = o.create_domain(2 * n).domainGen
        data = b"".join(o.fr_to_mont_bytes(x) for x in xs)
        res = out(32 * n)
        L.he_shift_eval(buf(data), res, lg)
        got = [o.fr_from_mont_bytes(bytes(res)[32 * i:32 * i + 32]) for i in range(n)]
        assert got == o.shift_eval_domain(xs, D, eta, fast=True), lg


def test_bucket_reduction_plan_emulation(L):
    """msm_reduce_plan.cuh (the plan MsmAccumulator::run executes): running sums, bit-sliced sums and the powers
    of two of the final assembly give sum_k (k+1) B_k for every bucket-set size the library can pick."""
    import ctypes as C
    import numpy as np
    M = (1 << 61) - 1
    L.he_reduce_emulate.restype = C.c_uint64
    rng = np.random.Generator(np.random.PCG64(11))
    seen_level2 = False
    for lg in list(range(0, 17)) + [19, 21]:
        nb = 1 << lg
        v = rng.integers(0, M, size=nb, dtype=np.uint64)
        info = (C.c_uint32 * 6)()
        got = L.he_reduce_emulate(v.ctypes.data_as(C.c_void_p), C.c_uint32(nb), info)
        k = np.arange(1, nb + 1, dtype=object)
        want = int(sum(int(a) * int(b) for a, b in zip(k, v.astype(object)))) % M if nb <= 1 << 12 else None
        if want is None:      # exact big-integer dot product in chunks
            want = 0
            vo = v.astype(object)
            for i in range(0, nb, 1 << 14):
                want += int(np.dot(np.arange(i + 1, min(nb, i + (1 << 14)) + 1, dtype=object), vo[i:i + (1 << 14)]))
            want %= M
        assert got == want, (lg, list(info))
        assert info[5] <= 4 and info[3] <= 20
        seen_level2 |= info[1] > 1
    assert seen_level2                      # the 2^19-bucket sets of the resident prover use level 2


def test_lazy_fp2_multiplication_model_is_carry_exact():
    """tools/emu_lazy_fp2.py: the instruction-level model of the lazily reduced Fp2 multiplication of field.cuh (wide
    products on two column-parity accumulators, reduction-only Montgomery loop, Karatsuba on unreduced values) -- every
    carry the PTX drops is asserted to be zero, results against Python integers, for both moduli and edge operands."""
    import os, random, sys
    tools = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    sys.path.insert(0, tools)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        import emu_lazy_fp2 as m
    rnd = random.Random(7)
    for mod in (m.P, m.R):
        inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
        rinv = pow(1 << 256, -1, mod)
        edge = [0, 1, mod - 1, mod - 2, (1 << 255) % mod, (1 << 224) - 1]
        for t in range(300):
            pick = (lambda: rnd.choice(edge)) if t < 100 else (lambda: rnd.randrange(mod))
            a0, a1, b0, b1 = pick(), pick(), pick(), pick()
            c0, c1 = m.fp2_mul_lazy(a0, a1, b0, b1, mod, inv)
            assert c0 == (a0 * b0 - a1 * b1) * rinv % mod and c1 == (a0 * b1 + a1 * b0) * rinv % mod
            x = rnd.randrange(mod << 256) if t >= 100 else rnd.choice([0, (mod << 256) - 1, (1 << 256) - 1, 1 << 256])
            assert m.redc(m.limbs(x, 16), mod, inv) == x * rinv % mod
        m.mul_wide((1 << 256) - 1, (1 << 256) - 1)
