"""Exact integer model of the FP64-pipe Montgomery multiply (52-bit limbs, R' = 2^260) used by field_fp64.cuh.

Models the bit patterns of the doubles produced by fma.rz.f64 and their accumulation with wrapping 64-bit adds,
including the pre-subtracted exponent offsets, and checks a*b*2^-260 mod p plus the limb bounds."""
import random

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
M52 = (1 << 52) - 1
M64 = (1 << 64) - 1
LO = 0x433 << 52     # exponent pattern of 2^52
HI = 0x467 << 52     # exponent pattern of 2^104
NL = [(P >> (52 * i)) & M52 for i in range(5)]
NP = (-pow(P, -1, 1 << 52)) % (1 << 52)


def split(a, b):
    """bit patterns of hi = fma_rz(a,b,2^104), lo = fma_rz(a,b,2^104+2^52-hi)"""
    assert 0 <= a < (1 << 52) and 0 <= b < (1 << 52)
    pr = a * b
    return HI + (pr >> 52), LO + (pr & M52)


def init_consts():
    # position k of the initial array (k = 0..4) is at position 0 in iteration k
    c = [(-(k * (2 * LO + 2 * HI) + 2 * LO)) & M64 for k in range(5)]
    # limbs born at position 5 at the start of iteration f end at result position f
    born = [(-(2 * HI + (4 - f) * (2 * LO + 2 * HI))) & M64 for f in range(5)]
    return c, born


def dfmul(a, b):
    c, born = init_consts()
    t = c + [born[0]]
    for i in range(5):
        for j in range(5):
            h, l = split(a[i], b[j])
            t[j] = (t[j] + l) & M64
            t[j + 1] = (t[j + 1] + h) & M64
        t0 = t[0] & M52
        _, ql = split(t0, NP)
        q = ql - LO
        for j in range(5):
            h, l = split(q, NL[j])
            t[j] = (t[j] + l) & M64
            t[j + 1] = (t[j + 1] + h) & M64
        assert t[0] & M52 == 0
        assert t[0] < (1 << 60), hex(t[0])            # offsets cancelled: a true small value
        carry = t[0] >> 52
        t = [(t[1] + carry) & M64, t[2], t[3], t[4], t[5], born[i + 1] if i < 4 else 0]
    for k in range(5):
        assert t[k] < (1 << 58), (k, hex(t[k]))
    for k in range(4):
        t[k + 1] += t[k] >> 52
        t[k] &= M52
    assert t[4] <= M52 and t[5] == 0
    return t[:5]


def limbs(x):
    return [(x >> (52 * i)) & M52 for i in range(5)]


def val(l):
    return sum(v << (52 * i) for i, v in enumerate(l))


if __name__ == "__main__":
    rnd = random.Random(1)
    Rinv = pow(1 << 260, -1, P)
    for it in range(20000):
        bound = rnd.choice([P, 2 * P, 9 * P, 1 << 256])
        a, b = rnd.randrange(bound), rnd.randrange(bound)
        if it < 4:
            a, b = [(0, 0), (P - 1, P - 1), (bound - 1, bound - 1), (1, 1)][it]
        r = val(dfmul(limbs(a), limbs(b)))
        assert r % P == a * b * Rinv % P
        assert r < a * b // (1 << 260) + P + 1
    print("ok; NL =", [hex(x) for x in NL], "NP =", hex(NP))
    print("consts", [hex(x) for x in init_consts()[0]], [hex(x) for x in init_consts()[1]])
