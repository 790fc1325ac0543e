"""CPU oracle for the Groth16 hot path of codex-storage/nim-groth16  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-Python (bigint) restatement of the reference's algorithms for the path
named by BASELINE.json:north_star.  It is the *checker*: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product (the CUDA
library under nim-groth16_b200/csrc) never calls into it.

PARITY UNPINNED: the reference (Nim + un-vendored mratsim/constantine @5f7ba18f) cannot be
compiled or run in this environment and its own tests hold no golden vectors
(tests/groth16/testProver.nim:65-73 only assert verifyProof == true under random toxic waste).
What pins this oracle instead: (1) all compared quantities are canonical (reduced field
elements, affine points, infinity = (0,0)) so any correct implementation is bit-identical;
(2) the constants of SURVEY.md Appendix B (re-derived in tests/test_oracle.py);
(3) the derived known-answer vectors of SURVEY.md Appendix C (tests/golden/);
(4) closed-form toxic-waste identities (fake setup => every MSM is (sum s_i k_i) * G);
(5) the pairing check of verifier.nim:31-52 restated in oracle/bn254_pairing.py.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------------------
# fields  (groth16/bn128/fields.nim:36-37)
# ----------------------------------------------------------------------------------------
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # fields.nim:36
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # fields.nim:37
MONT = 1 << 256                       # io.nim:60-65  (R = 2^256 Montgomery radix)
GEN28 = 0x2A3C09F0A58A7E8500E0A7EB8EF62ABC402D111E41112ED49BD61B6E725B19F0  # domain.nim:26
ONE_HALF_FR = 0x183227397098D014DC2822DB40C0AC2E9419F4243CDCB848A1F0FAC9F8000001  # ntt.nim:95


def inv_mod(a: int, m: int) -> int:
    return pow(a % m, -1, m)


# Fp2 = Fp[u]/(u^2+1)  (fields.nim:27,30-32; u^2 = -1 from export_sage.nim:92-95)
Fp2 = Tuple[int, int]


def fp2_add(a: Fp2, b: Fp2) -> Fp2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp2_sub(a: Fp2, b: Fp2) -> Fp2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fp2_mul(a: Fp2, b: Fp2) -> Fp2:
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fp2_sqr(a: Fp2) -> Fp2:
    return fp2_mul(a, a)


def fp2_neg(a: Fp2) -> Fp2:
    return ((-a[0]) % P, (-a[1]) % P)


def fp2_inv(a: Fp2) -> Fp2:
    d = inv_mod(a[0] * a[0] + a[1] * a[1], P)
    return (a[0] * d % P, (-a[1]) * d % P)


def fp2_scale(a: Fp2, k: int) -> Fp2:
    return (a[0] * k % P, a[1] * k % P)


FP2_ZERO: Fp2 = (0, 0)
FP2_ONE: Fp2 = (1, 0)

# ----------------------------------------------------------------------------------------
# curves  (groth16/bn128/curves.nim)
# ----------------------------------------------------------------------------------------
G1 = Tuple[int, int]          # affine, infinity = (0,0)      curves.nim:33,49
G2 = Tuple[Fp2, Fp2]          # affine, infinity = ((0,0),(0,0))  curves.nim:34,50
INF_G1: G1 = (0, 0)
INF_G2: G2 = (FP2_ZERO, FP2_ZERO)
GEN1: G1 = (1, 2)             # curves.nim:112-113
GEN2: G2 = (                  # curves.nim:115-121
    (0x1ADCD0ED10DF9CB87040F46655E3808F98AA68A570ACF5B0BDE23FAB1F149701,
     0x09E847E9F05A6082C3CD2A1D0A3A82E6FBFBE620F7F31269FA15D21C1C13B23B),
    (0x056C01168A5319461F7CA7AA19D4FCFD1C7CDF52DBFC4CBEE6F915250B7F6FC8,
     0x0EFE500A2D02DD77F5F401329F30895DF553B878FC3C0DADAAA86456A623235C),
)
TWIST_B: Fp2 = (              # curves.nim:75-77
    0x2B149D40CEB8AAAE81BE18991BE06AC3B5B4C5E559DBEFA33267E6DC24A138E5,
    0x009713B03AF0FED4CD2CAFADEED8FDF4A74FA084E52D1852E4A2BD0685C315D2,
)


def is_on_curve_g1(p: G1) -> bool:
    """curves.nim:54-67 checkCurveEqG1 (infinity counts as on-curve)."""
    x, y = p
    if x == 0 and y == 0:
        return True
    return (x * x * x + 3 - y * y) % P == 0


def is_on_curve_g2(p: G2) -> bool:
    """curves.nim:79-91 checkCurveEqG2."""
    x, y = p
    if x == FP2_ZERO and y == FP2_ZERO:
        return True
    return fp2_sub(fp2_add(fp2_mul(fp2_sqr(x), x), TWIST_B), fp2_sqr(y)) == FP2_ZERO


def g1_neg(p: G1) -> G1:
    return (p[0], (-p[1]) % P)


def g2_neg(p: G2) -> G2:
    return (p[0], fp2_neg(p[1]))


def g1_add(p: G1, q: G1) -> G1:
    """curves.nim:136-143 addG1 (affine -> projective sum -> affine); here the textbook
    affine chord/tangent law, which yields the same canonical affine point."""
    if p == INF_G1:
        return q
    if q == INF_G1:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF_G1
        lam = 3 * x1 * x1 * inv_mod(2 * y1, P) % P
    else:
        lam = (y2 - y1) * inv_mod(x2 - x1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g2_add(p: G2, q: G2) -> G2:
    """curves.nim:147-154 addG2."""
    if p == INF_G2:
        return q
    if q == INF_G2:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if fp2_add(y1, y2) == FP2_ZERO:
            return INF_G2
        lam = fp2_mul(fp2_scale(fp2_sqr(x1), 3), fp2_inv(fp2_scale(y1, 2)))
    else:
        lam = fp2_mul(fp2_sub(y2, y1), fp2_inv(fp2_sub(x2, x1)))
    x3 = fp2_sub(fp2_sub(fp2_sqr(lam), x1), x2)
    y3 = fp2_sub(fp2_mul(lam, fp2_sub(x1, x3)), y1)
    return (x3, y3)


# Jacobian helpers: only used to make scalar-mul / naive MSM fast enough in Python.
def _jac_dbl_g1(X, Y, Z):
    if Z == 0:
        return X, Y, Z
    A = X * X % P
    B = Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return X3, Y3, Z3


def _jac_madd_g1(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return x2, y2, 1
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    r = (S2 - Y1) % P
    if H == 0:
        if r == 0:
            return _jac_dbl_g1(x2, y2, 1)
        return 0, 1, 0
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (r * r - HHH - 2 * V) % P
    Y3 = (r * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return X3, Y3, Z3


def _jac_to_aff_g1(X, Y, Z) -> G1:
    if Z == 0:
        return INF_G1
    zi = inv_mod(Z, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_mul(k: int, p: G1) -> G1:
    """curves.nim:182-188 `**` (scalarMul_vartime on the projective lift, then affine)."""
    k %= R
    if p == INF_G1 or k == 0:
        return INF_G1
    X, Y, Z = 0, 1, 0
    for bit in bin(k)[2:]:
        X, Y, Z = _jac_dbl_g1(X, Y, Z)
        if bit == "1":
            X, Y, Z = _jac_madd_g1(X, Y, Z, p[0], p[1])
    return _jac_to_aff_g1(X, Y, Z)


def g2_mul(k: int, p: G2) -> G2:
    """curves.nim:190-196 `**` for G2 (plain double-and-add with affine adds)."""
    k %= R
    acc = INF_G2
    if p == INF_G2 or k == 0:
        return acc
    for bit in bin(k)[2:]:
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


# ----------------------------------------------------------------------------------------
# MSM  (groth16/bn128/msm.nim)
# ----------------------------------------------------------------------------------------
def msm_naive_g1(coeffs: Sequence[int], points: Sequence[G1]) -> G1:
    """msm.nim:162-178 msmNaiveG1: sum of coeffs[i] ** points[i]."""
    assert len(coeffs) == len(points), "incompatible sequence lengths"   # msm.nim:164
    X, Y, Z = 0, 1, 0
    for k, p in zip(coeffs, points):
        q = g1_mul(k, p)
        if q != INF_G1:
            X, Y, Z = _jac_madd_g1(X, Y, Z, q[0], q[1])
    return _jac_to_aff_g1(X, Y, Z)


def msm_naive_g2(coeffs: Sequence[int], points: Sequence[G2]) -> G2:
    """msm.nim:182-198 msmNaiveG2."""
    assert len(coeffs) == len(points), "incompatible sequence lengths"
    acc = INF_G2
    for k, p in zip(coeffs, points):
        acc = g2_add(acc, g2_mul(k, p))
    return acc


def msm_multithreaded_g1(nthreads_hint: int, coeffs: Sequence[int], points: Sequence[G1],
                         ncpu: int = 8) -> G1:
    """msm.nim:89-124 msmMultiThreadedG1: contiguous chunks [N*k/ntasks, N*(k+1)/ntasks),
    one MSM per chunk (msm.nim:35-59), affine partial sums added in order (msm.nim:117-119).
    The per-chunk MSM is any correct MSM (constantine's Pippenger, msm.nim:49); the result is
    canonical."""
    N = len(coeffs)
    assert N == len(points), "incompatible sequence lengths"             # msm.nim:97
    target = ncpu if nthreads_hint <= 0 else min(nthreads_hint, 256)     # msm.nim:98
    nthreads = max(1, min(N // 128, target))                             # msm.nim:99
    ntasks = nthreads if nthreads > 1 else 1                             # msm.nim:100
    res = INF_G1
    a = 0
    for k in range(ntasks):
        b = (N * (k + 1)) // ntasks if k < ntasks - 1 else N             # msm.nim:107-111
        res = g1_add(res, msm_naive_g1(coeffs[a:b], points[a:b]))
        a = b
    return res


def msm_multithreaded_g2(nthreads_hint: int, coeffs: Sequence[int], points: Sequence[G2],
                         ncpu: int = 8) -> G2:
    """msm.nim:128-158 msmMultiThreadedG2."""
    N = len(coeffs)
    assert N == len(points), "incompatible sequence lengths"
    target = ncpu if nthreads_hint <= 0 else min(nthreads_hint, 256)
    nthreads = max(1, min(N // 128, target))
    ntasks = nthreads if nthreads > 1 else 1
    res = INF_G2
    a = 0
    for k in range(ntasks):
        b = (N * (k + 1)) // ntasks if k < ntasks - 1 else N
        res = g2_add(res, msm_naive_g2(coeffs[a:b], points[a:b]))
        a = b
    return res


# ----------------------------------------------------------------------------------------
# domain  (groth16/math/domain.nim)
# ----------------------------------------------------------------------------------------
@dataclass
class Domain:                                   # domain.nim:16-21
    domainSize: int
    logDomainSize: int
    domainGen: int
    invDomainGen: int
    invDomainSize: int


def ceiling_log2(x: int) -> int:
    """misc.nim:43 ceilingLog2."""
    if x <= 1:
        return 0
    return (x - 1).bit_length()


def create_domain(size: int) -> Domain:
    """domain.nim:28-46 createDomain."""
    log2 = ceiling_log2(size)
    assert (1 << log2) == size, "domain must have a power-of-two size"   # domain.nim:30
    gen = pow(GEN28, 1 << (28 - log2), R)                                 # domain.nim:32-33
    assert pow(gen, size, R) == 1, "domain generator sanity check /A"     # domain.nim:38
    assert size == 1 or pow(gen, size // 2, R) != 1, "domain generator sanity check /B"
    return Domain(size, log2, gen, inv_mod(gen, R), inv_mod(size, R))


# ----------------------------------------------------------------------------------------
# NTT  (groth16/math/ntt.nim)  -- literal restatement of the recursive workers
# ----------------------------------------------------------------------------------------
def _forward_ntt_worker(m, src_stride, gpows, src, src_ofs, buf, buf_ofs, tgt, tgt_ofs):
    """ntt.nim:17-50 forwardNTT_worker (recursive DIT, strided source)."""
    if m == 0:
        tgt[tgt_ofs] = src[src_ofs]
    elif m == 1:
        tgt[tgt_ofs] = (src[src_ofs] + src[src_ofs + src_stride]) % R
        tgt[tgt_ofs + 1] = (src[src_ofs] - src[src_ofs + src_stride]) % R
    else:
        N = 1 << m
        half = 1 << (m - 1)
        _forward_ntt_worker(m - 1, src_stride << 1, gpows, src, src_ofs, buf, buf_ofs + N, buf, buf_ofs)
        _forward_ntt_worker(m - 1, src_stride << 1, gpows, src, src_ofs + src_stride, buf, buf_ofs + N,
                            buf, buf_ofs + half)
        for j in range(half):
            y = gpows[j * src_stride] * buf[buf_ofs + j + half] % R
            tgt[tgt_ofs + j] = (buf[buf_ofs + j] + y) % R
            tgt[tgt_ofs + j + half] = (buf[buf_ofs + j] - y) % R


def forward_ntt(src: Sequence[int], D: Domain) -> List[int]:
    """ntt.nim:55-77 forwardNTT: tgt[k] = sum_i src[i] * gen^(i k); natural order in/out."""
    assert D.domainSize == (1 << D.logDomainSize), "domain must have a power-of-two size"
    assert D.domainSize == len(src), "input must have the same size as the domain"
    N = D.domainSize
    buf = [0] * (2 * N)
    tgt = [0] * N
    gpows = [0] * (N // 2)
    x = 1
    for i in range(N // 2):                        # ntt.nim:64-69
        gpows[i] = x
        x = x * D.domainGen % R
    _forward_ntt_worker(D.logDomainSize, 1, gpows, list(src), 0, buf, 0, tgt, 0)
    return tgt


def _div2(x: int) -> int:
    return x * ONE_HALF_FR % R


def _inverse_ntt_worker(m, tgt_stride, gpows, src, src_ofs, buf, buf_ofs, tgt, tgt_ofs):
    """ntt.nim:97-135 inverseNTT_worker (recursive DIF, 1/2 folded into each level)."""
    if m == 0:
        tgt[tgt_ofs] = src[src_ofs]
    elif m == 1:
        tgt[tgt_ofs] = _div2((src[src_ofs] + src[src_ofs + 1]) % R)
        tgt[tgt_ofs + tgt_stride] = _div2((src[src_ofs] - src[src_ofs + 1]) % R)
    else:
        N = 1 << m
        half = 1 << (m - 1)
        for j in range(half):
            buf[buf_ofs + j] = _div2((src[src_ofs + j] + src[src_ofs + j + half]) % R)
            buf[buf_ofs + j + half] = (src[src_ofs + j] - src[src_ofs + j + half]) * gpows[j * tgt_stride] % R
        _inverse_ntt_worker(m - 1, tgt_stride << 1, gpows, buf, buf_ofs, buf, buf_ofs + N, tgt, tgt_ofs)
        _inverse_ntt_worker(m - 1, tgt_stride << 1, gpows, buf, buf_ofs + half, buf, buf_ofs + N, tgt,
                            tgt_ofs + tgt_stride)


def inverse_ntt(src: Sequence[int], D: Domain) -> List[int]:
    """ntt.nim:139-161 inverseNTT: exact inverse of forward_ntt, including the 1/N factor."""
    assert D.domainSize == (1 << D.logDomainSize), "domain must have a power-of-two size"
    assert D.domainSize == len(src), "input must have the same size as the domain"
    N = D.domainSize
    buf = [0] * (2 * N)
    tgt = [0] * N
    gpows = [0] * (N // 2)
    x = ONE_HALF_FR                                  # ntt.nim:149
    ginv = inv_mod(D.domainGen, R)
    for i in range(N // 2):
        gpows[i] = x
        x = x * ginv % R
    _inverse_ntt_worker(D.logDomainSize, 1, gpows, list(src), 0, buf, 0, tgt, 0)
    return tgt


def forward_ntt_fast(src: Sequence[int], D: Domain) -> List[int]:
    """Iterative radix-2 with the same definition as forward_ntt (for larger test sizes)."""
    n = D.domainSize
    lg = D.logDomainSize
    a = [0] * n
    for i in range(n):
        a[int(format(i, "0%db" % lg)[::-1], 2) if lg else 0] = src[i] % R
    half = 1
    while half < n:
        w = pow(D.domainGen, n // (2 * half), R)
        tw = [1] * half
        for j in range(1, half):
            tw[j] = tw[j - 1] * w % R
        for start in range(0, n, 2 * half):
            for j in range(half):
                u = a[start + j]
                v = a[start + j + half] * tw[j] % R
                a[start + j] = (u + v) % R
                a[start + j + half] = (u - v) % R
        half *= 2
    return a


def inverse_ntt_fast(src: Sequence[int], D: Domain) -> List[int]:
    Dinv = Domain(D.domainSize, D.logDomainSize, D.invDomainGen, D.domainGen, D.invDomainSize)
    out = forward_ntt_fast(src, Dinv)
    return [x * D.invDomainSize % R for x in out]


# ----------------------------------------------------------------------------------------
# zkey data model (groth16/zkey_types.nim)
# ----------------------------------------------------------------------------------------
JENS_GROTH = 0    # zkey_types.nim:11
SNARKJS = 1       # zkey_types.nim:12
MATRIX_A, MATRIX_B, MATRIX_C = 0, 1, 2     # zkey_types.nim:43-46


@dataclass
class Coeff:                    # zkey_types.nim:48-52
    matrix: int
    row: int
    col: int
    coeff: int


@dataclass
class ZKey:                     # zkey_types.nim:54-60 (header/spec/points flattened)
    flavour: int
    nvars: int
    npubs: int
    domainSize: int
    logDomainSize: int
    alpha1: G1
    beta1: G1
    beta2: G2
    gamma2: G2
    delta1: G1
    delta2: G2
    pointsIC: List[G1]
    pointsA1: List[G1]
    pointsB1: List[G1]
    pointsB2: List[G2]
    pointsC1: List[G1]
    pointsH1: List[G1]
    coeffs: List[Coeff]


@dataclass
class Proof:                    # prover.nim:38-43
    publicIO: List[int]
    pi_a: G1
    pi_b: G2
    pi_c: G1
    curve: str = "bn128"


# ----------------------------------------------------------------------------------------
# prover  (groth16/prover.nim)
# ----------------------------------------------------------------------------------------
def build_abc(zkey: ZKey, witness: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """prover.nim:56-73 buildABC.  C.w is never evaluated: Cz = Az o Bz (prover.nim:69-71);
    a matrix-C entry raises (prover.nim:67)."""
    n = zkey.domainSize
    Az = [0] * n
    Bz = [0] * n
    for e in zkey.coeffs:
        if e.matrix == MATRIX_A:
            Az[e.row] = (Az[e.row] + e.coeff * witness[e.col]) % R
        elif e.matrix == MATRIX_B:
            Bz[e.row] = (Bz[e.row] + e.coeff * witness[e.col]) % R
        else:
            raise AssertionError("fatal error")
    Cz = [Az[i] * Bz[i] % R for i in range(n)]
    return Az, Bz, Cz


def multiply_by_powers(xs: Sequence[int], eta: int) -> List[int]:
    """prover.nim:96-106 multiplyByPowers: ys[i] = eta^i * xs[i]."""
    n = len(xs)
    assert n >= 1
    ys = [0] * n
    ys[0] = xs[0]
    if n >= 1:
        ys[1] = eta * xs[1] % R        # prover.nim:101: indexes xs[1] => n = 1 is unsupported
    spow = eta
    for i in range(2, n):
        spow = spow * eta % R
        ys[i] = spow * xs[i] % R
    return ys


def shift_eval_domain(values: Sequence[int], D: Domain, eta: int, fast: bool = False) -> List[int]:
    """prover.nim:109-113 shiftEvalDomain: iNTT -> multiply by eta^i -> NTT."""
    intt = inverse_ntt_fast if fast else inverse_ntt
    fntt = forward_ntt_fast if fast else forward_ntt
    cs = intt(values, D)
    ds = multiply_by_powers(cs, eta)
    return fntt(ds, D)


def compute_snarkjs_scalar_coeffs(abc, fast: bool = False) -> List[int]:
    """prover.nim:158-181 computeSnarkjsScalarCoeffs: ys[j] = A1[j]*B1[j] - C1[j] on the coset."""
    Az, Bz, Cz = abc
    n = len(Az)
    assert len(Bz) == n and len(Cz) == n
    D = create_domain(n)
    eta = create_domain(2 * n).domainGen              # prover.nim:163
    A1 = shift_eval_domain(Az, D, eta, fast)
    B1 = shift_eval_domain(Bz, D, eta, fast)
    C1 = shift_eval_domain(Cz, D, eta, fast)
    return [(A1[j] * B1[j] - C1[j]) % R for j in range(n)]   # prover.nim:176


def compute_quotient_pointwise(abc, fast: bool = False) -> List[int]:
    """prover.nim:118-148 computeQuotientPointwise (JensGroth flavour): true quotient coeffs."""
    Az, Bz, Cz = abc
    n = len(Az)
    D = create_domain(n)
    eta = create_domain(2 * n).domainGen              # prover.nim:127
    invZ1 = inv_mod(pow(eta, n, R) - 1, R)            # prover.nim:128
    A1 = shift_eval_domain(Az, D, eta, fast)
    B1 = shift_eval_domain(Bz, D, eta, fast)
    C1 = shift_eval_domain(Cz, D, eta, fast)
    ys = [(A1[j] * B1[j] - C1[j]) * invZ1 % R for j in range(n)]   # prover.nim:141
    intt = inverse_ntt_fast if fast else inverse_ntt
    Q1 = intt(ys, D)                                   # prover.nim:142
    return multiply_by_powers(Q1, inv_mod(eta, R))     # prover.nim:143


def compute_qs(zkey: ZKey, abc, fast: bool = False) -> List[int]:
    """prover.nim:250-260 flavour switch."""
    if zkey.flavour == JENS_GROTH:
        return compute_quotient_pointwise(abc, fast)
    return compute_snarkjs_scalar_coeffs(abc, fast)


def generate_proof_with_mask(zkey: ZKey, witness: Sequence[int], r: int, s: int,
                             msm_g1=None, msm_g2=None, fast: bool = False,
                             intermediates: Optional[dict] = None) -> Proof:
    """prover.nim:215-304 generateProofWithMask."""
    msm_g1 = msm_g1 or msm_naive_g1
    msm_g2 = msm_g2 or msm_naive_g2
    nvars, npubs = zkey.nvars, zkey.npubs
    assert nvars == len(witness), "wrong witness length"                  # prover.nim:236
    pubIO = [witness[i] for i in range(npubs + 1)]                        # prover.nim:239-240
    abc = build_abc(zkey, witness)                                        # prover.nim:245
    qs = compute_qs(zkey, abc, fast)                                      # prover.nim:250-260
    zs = [witness[j] for j in range(npubs + 1, nvars)]                    # prover.nim:262-264
    assert len(witness) == len(zkey.pointsA1) == len(zkey.pointsB1) == len(zkey.pointsB2)
    assert zkey.domainSize == len(qs) == len(zkey.pointsH1)
    assert nvars - npubs - 1 == len(zs) == len(zkey.pointsC1)
    mA = msm_g1(witness, zkey.pointsA1)
    mB1 = msm_g1(witness, zkey.pointsB1)
    mB2 = msm_g2(witness, zkey.pointsB2)
    mH = msm_g1(qs, zkey.pointsH1)
    mC = msm_g1(zs, zkey.pointsC1)
    pi_a = g1_add(g1_add(zkey.alpha1, g1_mul(r, zkey.delta1)), mA)        # prover.nim:279-282
    rho = g1_add(g1_add(zkey.beta1, g1_mul(s, zkey.delta1)), mB1)         # prover.nim:285-288
    pi_b = g2_add(g2_add(zkey.beta2, g2_mul(s, zkey.delta2)), mB2)        # prover.nim:291-294
    pi_c = g1_mul(s, pi_a)                                                # prover.nim:298
    pi_c = g1_add(pi_c, g1_mul(r, rho))                                   # prover.nim:299
    pi_c = g1_add(pi_c, g1_mul((-(r * s)) % R, zkey.delta1))              # prover.nim:300
    pi_c = g1_add(pi_c, mH)                                               # prover.nim:301
    pi_c = g1_add(pi_c, mC)                                               # prover.nim:302
    if intermediates is not None:
        intermediates.update(Az=abc[0], Bz=abc[1], Cz=abc[2], qs=qs, msmA=mA, msmB1=mB1, msmB2=mB2,
                             msmH=mH, msmC=mC, rho=rho)
    return Proof(publicIO=pubIO, pi_a=pi_a, pi_b=pi_b, pi_c=pi_c)


# ----------------------------------------------------------------------------------------
# R1CS + fake setup  (groth16/files/r1cs.nim:64-80, groth16/fake_setup.nim)
# ----------------------------------------------------------------------------------------
Term = Tuple[int, int]                       # (wireIdx, value)         r1cs.nim:72
Constraint = Tuple[List[Term], List[Term], List[Term]]   # (A, B, C)   r1cs.nim:74


@dataclass
class R1CS:                                  # r1cs.nim:64-80
    nWires: int
    nPubOut: int
    nPubIn: int
    nPrivIn: int
    constraints: List[Constraint]
    nLabels: int = 0
    wireToLabel: List[int] = field(default_factory=list)


@dataclass
class ToxicWaste:                            # fake_setup.nim:24-30
    alpha: int
    beta: int
    gamma: int
    delta: int
    tau: int


def r1cs_to_coeffs(r1cs: R1CS) -> List[Coeff]:
    """fake_setup.nim:46-66 r1csToCoeffs, including the npub+1 dummy rows A[n+i][i] = 1."""
    coeffs: List[Coeff] = []
    n = len(r1cs.constraints)
    p = r1cs.nPubIn + r1cs.nPubOut
    for i, (A, B, _C) in enumerate(r1cs.constraints):
        for (w, v) in A:
            coeffs.append(Coeff(MATRIX_A, i, w, v % R))
        for (w, v) in B:
            coeffs.append(Coeff(MATRIX_B, i, w, v % R))
    for i in range(n, n + p + 1):                       # fake_setup.nim:61-63
        coeffs.append(Coeff(MATRIX_A, i, i - n, 1))
    return coeffs


def eval_lagrange_poly_at(D: Domain, k: int, zeta: int) -> int:
    """poly.nim:242-250 evalLagrangePolyAt."""
    omega_k = pow(D.domainGen, k, R)
    denom = (zeta - omega_k) % R
    if denom == 0:
        raise AssertionError("point should be outside the domain")
    return omega_k * (pow(zeta, D.domainSize, R) - 1) % R * D.invDomainSize % R * inv_mod(denom, R) % R


def _batch_lagrange(D: Domain, ks: Sequence[int], zeta: int) -> List[int]:
    """Same values as eval_lagrange_poly_at for many k, with one batch inversion."""
    zn1 = (pow(zeta, D.domainSize, R) - 1) % R
    oms = [pow(D.domainGen, k, R) for k in ks] if len(ks) < 64 else None
    if oms is None:
        # consecutive-stride fast path
        oms = []
        if len(ks) >= 2:
            step = pow(D.domainGen, ks[1] - ks[0], R)
        else:
            step = 1
        x = pow(D.domainGen, ks[0], R)
        for _ in ks:
            oms.append(x)
            x = x * step % R
    den = [(zeta - o) % R for o in oms]
    pref = [1] * (len(den) + 1)
    for i, d in enumerate(den):
        assert d != 0, "point should be outside the domain"
        pref[i + 1] = pref[i] * d % R
    inv_all = inv_mod(pref[-1], R)
    out = [0] * len(den)
    for i in range(len(den) - 1, -1, -1):
        out[i] = inv_all * pref[i] % R
        inv_all = inv_all * den[i] % R
    c = zn1 * D.invDomainSize % R
    return [oms[i] * c % R * out[i] % R for i in range(len(den))]


@dataclass
class SetupScalars:
    """Discrete logs of every zkey point under a fake setup (closed-form oracle)."""
    a: List[int]
    b: List[int]
    c: List[int]
    ic: List[int]
    k: List[int]
    h: List[int]


def fake_setup_scalars(r1cs: R1CS, toxic: ToxicWaste, flavour: int = SNARKJS) -> Tuple[SetupScalars, int, int]:
    """fake_setup.nim:201-304: the field-side half of fakeCircuitSetup."""
    neqs = len(r1cs.constraints)
    npub = r1cs.nPubIn + r1cs.nPubOut
    logn = ceiling_log2(neqs + npub + 1)               # fake_setup.nim:205
    n = 1 << logn
    nvars = r1cs.nWires
    D = create_domain(n)
    L = _batch_lagrange(D, list(range(n)), toxic.tau)  # fake_setup.nim:255
    a = [0] * nvars
    b = [0] * nvars
    c = [0] * nvars
    for i, (A, B, C) in enumerate(r1cs.constraints):   # fake_setup.nim:159-187 + 264-266
        for (w, v) in A:
            a[w] = (a[w] + v * L[i]) % R
        for (w, v) in B:
            b[w] = (b[w] + v * L[i]) % R
        for (w, v) in C:
            c[w] = (c[w] + v * L[i]) % R
    for i in range(neqs, neqs + npub + 1):             # fake_setup.nim:182-185
        a[i - neqs] = (a[i - neqs] + L[i]) % R
    ginv = inv_mod(toxic.gamma, R)
    dinv = inv_mod(toxic.delta, R)
    comb = [(toxic.beta * a[j] + toxic.alpha * b[j] + c[j]) % R for j in range(nvars)]
    ic = [ginv * comb[j] % R for j in range(npub + 1)]                  # fake_setup.nim:276-277
    k = [dinv * comb[j] % R for j in range(npub + 1, nvars)]            # fake_setup.nim:279-280
    if flavour == JENS_GROTH:                                            # fake_setup.nim:293-295
        ztau = (pow(toxic.tau, n, R) - 1) % R
        h = []
        x = dinv * ztau % R
        for _ in range(n):
            h.append(x)
            x = x * toxic.tau % R
    else:                                                                # fake_setup.nim:301-304
        D2 = create_domain(2 * n)
        L2 = _batch_lagrange(D2, [2 * i + 1 for i in range(n)], toxic.tau)
        h = [dinv * x % R for x in L2]
    return SetupScalars(a, b, c, ic, k, h), n, logn


def fake_circuit_setup(r1cs: R1CS, toxic: ToxicWaste, flavour: int = SNARKJS,
                       g1_fixed=None, g2_fixed=None) -> Tuple[ZKey, SetupScalars]:
    """fake_setup.nim:201-326 fakeCircuitSetup with explicit toxic waste.
    g1_fixed / g2_fixed: optional batch fixed-base multipliers (list[int] -> list[point])."""
    sc, n, logn = fake_setup_scalars(r1cs, toxic, flavour)
    mul1 = g1_fixed or (lambda ks: [g1_mul(k, GEN1) for k in ks])
    mul2 = g2_fixed or (lambda ks: [g2_mul(k, GEN2) for k in ks])
    npub = r1cs.nPubIn + r1cs.nPubOut
    zkey = ZKey(
        flavour=flavour, nvars=r1cs.nWires, npubs=npub, domainSize=n, logDomainSize=logn,
        alpha1=g1_mul(toxic.alpha, GEN1), beta1=g1_mul(toxic.beta, GEN1),
        beta2=g2_mul(toxic.beta, GEN2), gamma2=g2_mul(toxic.gamma, GEN2),
        delta1=g1_mul(toxic.delta, GEN1), delta2=g2_mul(toxic.delta, GEN2),
        pointsIC=mul1(sc.ic), pointsA1=mul1(sc.a), pointsB1=mul1(sc.b), pointsB2=mul2(sc.b),
        pointsC1=mul1(sc.k), pointsH1=mul1(sc.h), coeffs=r1cs_to_coeffs(r1cs))
    return zkey, sc


def closed_form_proof_scalars(sc: SetupScalars, toxic: ToxicWaste, npubs: int,
                              witness: Sequence[int], qs: Sequence[int], r: int, s: int):
    """discrete logs (a, b, c) of (pi_a, pi_b, pi_c) and of the five MSMs."""
    dot = lambda xs, ys: sum(x * y for x, y in zip(xs, ys)) % R
    mA = dot(witness, sc.a)
    mB = dot(witness, sc.b)
    mH = dot(qs, sc.h)
    mC = dot(witness[npubs + 1:], sc.k)
    a = (toxic.alpha + r * toxic.delta + mA) % R
    b = (toxic.beta + s * toxic.delta + mB) % R
    c = (s * a + r * b - r * s * toxic.delta + mH + mC) % R
    return dict(a=a, b=b, c=c, msmA=mA, msmB=mB, msmH=mH, msmC=mC)


def closed_form_check(sc: SetupScalars, toxic: ToxicWaste, npubs: int, witness, cf) -> bool:
    """verifier.nim:31-52 in the exponent: a*b == alpha*beta + pub*gamma + c*delta."""
    pub = sum(witness[j] * sc.ic[j] for j in range(npubs + 1)) % R
    return (cf["a"] * cf["b"] - toxic.alpha * toxic.beta - pub * toxic.gamma - cf["c"] * toxic.delta) % R == 0


# ----------------------------------------------------------------------------------------
# the reference test circuit (tests/groth16/testProver.nim:17-47)
# ----------------------------------------------------------------------------------------
def reference_test_r1cs() -> R1CS:
    m1 = R - 1
    eq1 = ([], [], [(1, m1), (2, 1), (7, 1)])          # testProver.nim:26
    eq2 = ([(3, 1)], [(4, 1)], [(6, 1)])               # testProver.nim:29
    eq3 = ([(5, 1)], [(6, 1)], [(7, 1)])               # testProver.nim:32
    return R1CS(nWires=8, nPubOut=1, nPubIn=1, nPrivIn=3, constraints=[eq1, eq2, eq3])


REFERENCE_TEST_WITNESS = [1, 2023, 1022, 7, 11, 13, 77, 1001]   # testProver.nim:47


# ----------------------------------------------------------------------------------------
# synthetic chain circuit: (x_j + c_j) * x_j = x_{j+1}, last: x_last * 1 = out
# ----------------------------------------------------------------------------------------
def splitmix64(state: int) -> Tuple[int, int]:
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


class Rng:
    """Deterministic PRNG for fixtures (splitmix64 stream; Fr = 4 x u64 reduced mod r)."""

    def __init__(self, seed: int):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def u64(self) -> int:
        self.state, out = splitmix64(self.state)
        return out

    def fr(self) -> int:
        v = 0
        for i in range(4):
            v |= self.u64() << (64 * i)
        return v % R


def synthetic_r1cs(neqs: int, seed: int = 3) -> Tuple[R1CS, List[int]]:
    """nPubOut = 1, nPubIn = 0; wires 0:1, 1:out, 2:x0, 3..:x_{j+1}; returns (r1cs, witness)."""
    rng = Rng(seed)
    cons: List[Constraint] = []
    nw = neqs + 2
    wit = [0] * nw
    wit[0] = 1
    x = rng.fr()
    wit[2] = x
    for j in range(neqs - 1):
        c = rng.fr()
        cons.append(([(2 + j, 1), (0, c)], [(2 + j, 1)], [(3 + j, 1)]))
        x = (x + c) * x % R
        wit[3 + j] = x
    cons.append(([(neqs + 1, 1)], [(0, 1)], [(1, 1)]))
    wit[1] = wit[neqs + 1]
    return R1CS(nWires=nw, nPubOut=1, nPubIn=0, nPrivIn=1, constraints=cons), wit


# ----------------------------------------------------------------------------------------
# byte encodings (groth16/bn128/io.nim:103-153) and file formats
# ----------------------------------------------------------------------------------------
def fr_to_std_bytes(x: int) -> bytes:           # io.nim:141-145 (.wtns / .r1cs encoding)
    return (x % R).to_bytes(32, "little")


def fr_to_mont_bytes(x: int) -> bytes:          # in-memory constantine Fr / io.nim:147-152
    return (x * MONT % R).to_bytes(32, "little")


def fr_to_wtf_bytes(x: int) -> bytes:           # io.nim:134-139 (.zkey coefficients: R^2)
    return (x * MONT % R * MONT % R).to_bytes(32, "little")


def fp_to_mont_bytes(x: int) -> bytes:          # io.nim:126-131 (.zkey points)
    return (x * MONT % P).to_bytes(32, "little")


def fr_from_std_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * inv_mod(MONT, R) % R


def fp_from_mont_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little") * inv_mod(MONT, P) % P


def g1_to_bytes(p: G1) -> bytes:                # io.nim:228-231 layout: x then y
    return fp_to_mont_bytes(p[0]) + fp_to_mont_bytes(p[1])


def g2_to_bytes(p: G2) -> bytes:                # io.nim:198-201,233-236: x.c0 x.c1 y.c0 y.c1
    return b"".join(fp_to_mont_bytes(v) for v in (p[0][0], p[0][1], p[1][0], p[1][1]))


def g1_from_bytes(b: bytes) -> G1:
    return (fp_from_mont_bytes(b[0:32]), fp_from_mont_bytes(b[32:64]))


def g2_from_bytes(b: bytes) -> G2:
    v = [fp_from_mont_bytes(b[32 * i:32 * i + 32]) for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


def _container(magic: bytes, version: int, sections: Sequence[Tuple[int, bytes]]) -> bytes:
    """files/container.nim:1-20 iden3 binfile: magic, version, nsections, (id u32, len u64, data)*."""
    out = [magic, struct.pack("<II", version, len(sections))]
    for sid, data in sections:
        out.append(struct.pack("<IQ", sid, len(data)))
        out.append(data)
    return b"".join(out)


def write_zkey_bytes(zkey: ZKey) -> bytes:
    """files/zkey.nim:1-92 layout (sections 1..9); the reference has no writer."""
    s1 = struct.pack("<I", 1)
    s2 = (struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little")
          + struct.pack("<III", zkey.nvars, zkey.npubs, zkey.domainSize)
          + g1_to_bytes(zkey.alpha1) + g1_to_bytes(zkey.beta1) + g2_to_bytes(zkey.beta2)
          + g2_to_bytes(zkey.gamma2) + g1_to_bytes(zkey.delta1) + g2_to_bytes(zkey.delta2))
    s3 = b"".join(g1_to_bytes(p) for p in zkey.pointsIC)
    s4 = struct.pack("<I", len(zkey.coeffs)) + b"".join(
        struct.pack("<III", c.matrix, c.row, c.col) + fr_to_wtf_bytes(c.coeff) for c in zkey.coeffs)
    s5 = b"".join(g1_to_bytes(p) for p in zkey.pointsA1)
    s6 = b"".join(g1_to_bytes(p) for p in zkey.pointsB1)
    s7 = b"".join(g2_to_bytes(p) for p in zkey.pointsB2)
    s8 = b"".join(g1_to_bytes(p) for p in zkey.pointsC1)
    s9 = b"".join(g1_to_bytes(p) for p in zkey.pointsH1)
    return _container(b"zkey", 1, [(1, s1), (2, s2), (3, s3), (4, s4), (5, s5), (6, s6), (7, s7), (8, s8), (9, s9)])


def write_wtns_bytes(witness: Sequence[int]) -> bytes:
    """files/witness.nim:1-15,36-60 layout."""
    s1 = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", len(witness))
    s2 = b"".join(fr_to_std_bytes(x) for x in witness)
    return _container(b"wtns", 2, [(1, s1), (2, s2)])


def write_r1cs_bytes(r1cs: R1CS) -> bytes:
    """files/r1cs.nim:1-50 layout."""
    s1 = (struct.pack("<I", 32) + R.to_bytes(32, "little")
          + struct.pack("<IIIIQI", r1cs.nWires, r1cs.nPubOut, r1cs.nPubIn, r1cs.nPrivIn, r1cs.nLabels,
                        len(r1cs.constraints)))
    parts = []
    for con in r1cs.constraints:
        for lc in con:
            parts.append(struct.pack("<I", len(lc)))
            for (w, v) in lc:
                parts.append(struct.pack("<I", w) + fr_to_std_bytes(v))
    s2 = b"".join(parts)
    s3 = b"".join(struct.pack("<Q", i) for i in range(r1cs.nWires))
    return _container(b"r1cs", 1, [(1, s1), (2, s2), (3, s3)])


def parse_container(data: bytes, magic: bytes, version: int) -> dict:
    """files/container.nim:75-93 parseContainer."""
    assert data[0:4] == magic, "not a `%s` file" % magic.decode()
    ver, nsec = struct.unpack_from("<II", data, 4)
    assert ver == version
    pos = 12
    sections = {}
    for _ in range(nsec):
        sid, slen = struct.unpack_from("<IQ", data, pos)
        pos += 12
        sections[sid] = data[pos:pos + slen]
        pos += slen
    return sections


def parse_zkey_bytes(data: bytes) -> ZKey:
    """files/zkey.nim:114-248 parseZKey (flavour hard-coded Snarkjs, zkey.nim:129)."""
    sec = parse_container(data, b"zkey", 1)
    assert struct.unpack("<I", sec[1])[0] == 1, "expecting `.zkey` file for a Groth16 prover"
    s2 = sec[2]
    assert struct.unpack_from("<I", s2, 0)[0] == 32 and int.from_bytes(s2[4:36], "little") == P
    assert struct.unpack_from("<I", s2, 36)[0] == 32 and int.from_bytes(s2[40:72], "little") == R
    nvars, npubs, dom = struct.unpack_from("<III", s2, 72)
    o = 84
    alpha1 = g1_from_bytes(s2[o:o + 64]); o += 64
    beta1 = g1_from_bytes(s2[o:o + 64]); o += 64
    beta2 = g2_from_bytes(s2[o:o + 128]); o += 128
    gamma2 = g2_from_bytes(s2[o:o + 128]); o += 128
    delta1 = g1_from_bytes(s2[o:o + 64]); o += 64
    delta2 = g2_from_bytes(s2[o:o + 128]); o += 128
    g1s = lambda b: [g1_from_bytes(b[i:i + 64]) for i in range(0, len(b), 64)]
    g2s = lambda b: [g2_from_bytes(b[i:i + 128]) for i in range(0, len(b), 128)]
    nco = struct.unpack_from("<I", sec[4], 0)[0]
    assert len(sec[4]) == 4 + nco * 44, "unexpected section length"
    rinv2 = inv_mod(MONT, R) ** 2 % R
    coeffs = []
    for i in range(nco):
        m, rr, cc = struct.unpack_from("<III", sec[4], 4 + 44 * i)
        v = int.from_bytes(sec[4][4 + 44 * i + 12:4 + 44 * i + 44], "little") * rinv2 % R
        coeffs.append(Coeff(m, rr, cc, v))
    return ZKey(flavour=SNARKJS, nvars=nvars, npubs=npubs, domainSize=dom, logDomainSize=ceiling_log2(dom),
                alpha1=alpha1, beta1=beta1, beta2=beta2, gamma2=gamma2, delta1=delta1, delta2=delta2,
                pointsIC=g1s(sec[3]), pointsA1=g1s(sec[5]), pointsB1=g1s(sec[6]), pointsB2=g2s(sec[7]),
                pointsC1=g1s(sec[8]), pointsH1=g1s(sec[9]), coeffs=coeffs)


def parse_wtns_bytes(data: bytes) -> List[int]:
    """files/witness.nim:36-76 parseWitness."""
    sec = parse_container(data, b"wtns", 2)
    assert struct.unpack_from("<I", sec[1], 0)[0] == 32
    assert int.from_bytes(sec[1][4:36], "little") == R, "expecting the alt-bn128 curve"
    nvars = struct.unpack_from("<I", sec[1], 36)[0]
    assert len(sec[2]) == 32 * nvars
    return [int.from_bytes(sec[2][32 * i:32 * i + 32], "little") for i in range(nvars)]
