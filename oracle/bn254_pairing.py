"""BN254 optimal-ate pairing and the Groth16 verifier -- TEST INFRASTRUCTURE ONLY (part of the oracle).

Restates groth16/verifier.nim:31-52 (verifyProof) and the `pairing` wrapper of groth16/bn128/curves.nim:218-221.
The reference delegates the pairing to constantine (`pairing_bn[BN254Snarks]`, not vendored), so the published
algorithm is restated here in plain Python bigints: Fp12 = Fp[w]/(w^12 - 18 w^6 + 82) (i.e. Fp2[w]/(w^6 - (9+u))
flattened), G2 points untwisted into E(Fp12), Miller loop over 6x+2 with x = 4965661367192848881, two Frobenius
correction steps, final exponentiation (p^12 - 1)/r.  GT values are canonical, so e(P,Q) agrees with any other
correct implementation; the checks that pin it are bilinearity, non-degeneracy and the verifier equation on
proofs whose validity is known in the exponent (tests/test_oracle_pairing.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from g16_oracle import GEN1, GEN2, INF_G1, INF_G2, P, R, g1_add, g1_mul, g1_neg, is_on_curve_g1, is_on_curve_g2

ATE_LOOP_COUNT = 29793968203157093288          # 6x + 2
LOG_ATE_LOOP_COUNT = 63
FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]   # w^12 = 18 w^6 - 82


class F12:
    """Element of Fp[w]/(w^12 - 18 w^6 + 82), coefficients low degree first."""
    __slots__ = ("c",)

    def __init__(self, coeffs: Sequence[int]):
        self.c = [x % P for x in coeffs]

    @staticmethod
    def one() -> "F12":
        return F12([1] + [0] * 11)

    @staticmethod
    def zero() -> "F12":
        return F12([0] * 12)

    def __add__(self, o: "F12") -> "F12":
        return F12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o: "F12") -> "F12":
        return F12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self) -> "F12":
        return F12([-a for a in self.c])

    def scale(self, k: int) -> "F12":
        return F12([a * k for a in self.c])

    def __mul__(self, o: "F12") -> "F12":
        t = [0] * 23
        a, b = self.c, o.c
        for i in range(12):
            ai = a[i]
            if ai:
                for j in range(12):
                    t[i + j] += ai * b[j]
        for k in range(22, 11, -1):          # w^k = 18 w^(k-6) - 82 w^(k-12)
            v = t[k]
            if v:
                t[k - 6] += 18 * v
                t[k - 12] -= 82 * v
        return F12(t[:12])

    def __eq__(self, o) -> bool:
        return self.c == o.c

    def is_zero(self) -> bool:
        return not any(self.c)

    def inv(self) -> "F12":
        """Extended Euclid on polynomials over Fp."""
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [x % P for x in FQ12_MOD] + [1]
        deg = lambda p: max([i for i, v in enumerate(p) if v] or [0])
        while deg(low):
            # r = high // low (polynomial rounded division)
            dl, dh = deg(low), deg(high)
            temp = list(high)
            out = [0] * 13
            inv_lead = pow(low[dl], -1, P)
            for i in range(dh - dl, -1, -1):
                q = temp[dl + i] * inv_lead % P
                out[i] = q
                if q:
                    for c in range(dl + 1):
                        temp[c + i] = (temp[c + i] - low[c] * q) % P
            nm, new = list(hm), list(high)
            for i in range(13):
                if lm[i] or low[i]:
                    for j in range(13 - i):
                        if out[j]:
                            nm[i + j] = (nm[i + j] - lm[i] * out[j]) % P
                            new[i + j] = (new[i + j] - low[i] * out[j]) % P
            lm, low, hm, high = nm, new, lm, low
        k = pow(low[0], -1, P)
        return F12([x * k for x in lm[:12]])

    def __truediv__(self, o: "F12") -> "F12":
        return self * o.inv()

    def pow(self, e: int) -> "F12":
        res, base = F12.one(), self
        while e:
            if e & 1:
                res = res * base
            base = base * base
            e >>= 1
        return res


W = F12([0, 1] + [0] * 10)
W2, W3 = W * W, W * W * W
Pt12 = Tuple[F12, F12]


def _embed_fp(x: int) -> F12:
    return F12([x] + [0] * 11)


def twist(q) -> Pt12:
    """E'(Fp2) -> E(Fp12): (x, y) -> (x w^2, y w^3) with u = w^6 - 9."""
    (x0, x1), (y0, y1) = q
    nx = F12([x0 - 9 * x1, 0, 0, 0, 0, 0, x1, 0, 0, 0, 0, 0])
    ny = F12([y0 - 9 * y1, 0, 0, 0, 0, 0, y1, 0, 0, 0, 0, 0])
    return (nx * W2, ny * W3)


def _dbl(p: Pt12) -> Pt12:
    x, y = p
    lam = (x * x).scale(3) / y.scale(2)
    nx = lam * lam - x.scale(2)
    return (nx, lam * (x - nx) - y)


def _add(p: Pt12, q: Pt12) -> Pt12:
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        assert y1 == y2, "P + (-P) does not occur in the Miller loop of a valid input"
        return _dbl(p)
    lam = (y2 - y1) / (x2 - x1)
    nx = lam * lam - x1 - x2
    return (nx, lam * (x1 - nx) - y1)


def _line(p1: Pt12, p2: Pt12, t: Pt12) -> F12:
    (x1, y1), (x2, y2), (xt, yt) = p1, p2, t
    if not (x1 == x2):
        lam = (y2 - y1) / (x2 - x1)
        return lam * (xt - x1) - (yt - y1)
    if y1 == y2:
        lam = (x1 * x1).scale(3) / y1.scale(2)
        return lam * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q, p) -> F12:
    """Miller function f_{6x+2,Q}(P) with the two Frobenius lines; no final exponentiation."""
    if p == INF_G1 or q == INF_G2:
        return F12.one()
    Q = twist(q)
    Pp = (_embed_fp(p[0]), _embed_fp(p[1]))
    Rr, f = Q, F12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _line(Rr, Rr, Pp)
        Rr = _dbl(Rr)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _line(Rr, Q, Pp)
            Rr = _add(Rr, Q)
    Q1 = (Q[0].pow(P), Q[1].pow(P))
    nQ2 = (Q1[0].pow(P), -(Q1[1].pow(P)))
    f = f * _line(Rr, Q1, Pp)
    Rr = _add(Rr, Q1)
    f = f * _line(Rr, nQ2, Pp)
    return f


def final_exponentiate(f: F12) -> F12:
    return f.pow((P ** 12 - 1) // R)


def pairing(p, q) -> F12:
    """curves.nim:218-221 pairing(p: G1, q: G2)."""
    assert is_on_curve_g1(p) and is_on_curve_g2(q)
    return final_exponentiate(miller_loop(q, p))


def msm_small_g1(coeffs: Sequence[int], points) -> Tuple[int, int]:
    acc = INF_G1
    for k, pt in zip(coeffs, points):
        acc = g1_add(acc, g1_mul(k, pt))
    return acc


def verify_proof(alpha1, beta2, gamma2, delta2, points_ic, public_io: Sequence[int], pi_a, pi_b, pi_c) -> bool:
    """verifier.nim:31-52: e(-pi_a, pi_b) * e(alpha1, beta2) * e(pi_c, delta2) * e(sum pub_j IC_j, gamma2) == 1.
    (The reference multiplies by the stored alphaBeta = e(alpha1, beta2), zkey.nim:164.)  One shared final
    exponentiation over the product of the four Miller functions."""
    assert is_on_curve_g1(pi_a), "pi_a is not in G1"            # verifier.nim:35
    assert is_on_curve_g2(pi_b), "pi_b is not in G2"            # verifier.nim:36
    assert is_on_curve_g1(pi_c), "pi_c is not in G1"            # verifier.nim:37
    assert len(public_io) == len(points_ic)
    pub = msm_small_g1(public_io, points_ic)                    # verifier.nim:39
    f = miller_loop(pi_b, g1_neg(pi_a))                         # verifier.nim:41
    f = f * miller_loop(beta2, alpha1)                          # verifier.nim:42
    f = f * miller_loop(delta2, pi_c)                           # verifier.nim:43
    f = f * miller_loop(gamma2, pub)                            # verifier.nim:44
    return final_exponentiate(f) == F12.one()                   # verifier.nim:52
