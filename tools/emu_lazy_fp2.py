"""Carry-exact model of the lazily reduced Fp2 multiplication of csrc/field.cuh: 8x8 -> 16 limb products on two
column-parity accumulators (rows of lo/hi multiply-add pairs that ptxas fuses into IMAD.WIDE), Montgomery reduction
of a 16-limb value with the even/odd shifting accumulators of fmul, Karatsuba over three unreduced products and two
reductions instead of three.  Run: python tools/emu_lazy_fp2.py"""
import random
from emu_montmul import CC, limbs, val, M32, P, R


def row_carry(cc, Y, off, xs, b, carry_into=True):
    """Y[off..off+7] += {xs[0..3]} * b as four lo/hi pairs; the carry-out goes to Y[off+8] (or must be zero)."""
    Y[off + 0] = cc.mad_lo_cc(xs[0], b, Y[off + 0]); Y[off + 1] = cc.madc_hi_cc(xs[0], b, Y[off + 1])
    Y[off + 2] = cc.madc_lo_cc(xs[1], b, Y[off + 2]); Y[off + 3] = cc.madc_hi_cc(xs[1], b, Y[off + 3])
    Y[off + 4] = cc.madc_lo_cc(xs[2], b, Y[off + 4]); Y[off + 5] = cc.madc_hi_cc(xs[2], b, Y[off + 5])
    Y[off + 6] = cc.madc_lo_cc(xs[3], b, Y[off + 6])
    if carry_into:
        Y[off + 7] = cc.madc_hi_cc(xs[3], b, Y[off + 7])
        Y[off + 8] = cc.addc(Y[off + 8], 0)
    else:
        Y[off + 7] = cc.madc_hi(xs[3], b, Y[off + 7])      # asserts: no carry-out


def mul_wide(a, b):
    """E[k] sits at column k, O[k] at column k+1; product = E + O * 2^32."""
    cc = CC(); A = limbs(a); B = limbs(b)
    ev, od = [A[0], A[2], A[4], A[6]], [A[1], A[3], A[5], A[7]]
    E = [0] * 17; O = [0] * 17
    for i in range(8):
        if i % 2 == 0:
            row_carry(cc, E, i, ev, B[i])
            row_carry(cc, O, i, od, B[i])
        else:
            row_carry(cc, O, i - 1, ev, B[i])
            row_carry(cc, E, i + 1, od, B[i], carry_into=(i < 7))
    assert E[16] == 0 and O[15] == 0 and O[16] == 0
    r = [0] * 16
    r[0] = E[0]
    r[1] = cc.add_cc(E[1], O[0])
    for k in range(2, 15):
        r[k] = cc.addc_cc(E[k], O[k - 1])
    r[15] = cc.addc(E[15], O[14])
    assert val(r) == a * b
    return r


def redc(T, p, inv):
    """T: 16 limbs, value < p * 2^256.  Returns T / 2^256 mod p (fully reduced)."""
    cc = CC(); Pm = limbs(p)
    pe, po = [Pm[0], Pm[2], Pm[4], Pm[6]], [Pm[1], Pm[3], Pm[5], Pm[7]]
    t = [T[0:8] + [0], [0] * 9]            # t[0] even role first (holds T_lo), t[1] odd role; slot 8 = guard (unused)
    for i in range(8):
        Y = t[i & 1]; X = t[(i + 1) & 1]
        if i > 0:
            # shift: the previous even array (now X) has limb 0 == 0; its limb 1 joins column 0 of Y, the rest moves down
            Y[0] = cc.add_cc(Y[0], X[1])
            X[0] = cc.addc_cc(X[2], 0); X[1] = cc.addc_cc(X[3], 0); X[2] = cc.addc_cc(X[4], 0)
            X[3] = cc.addc_cc(X[5], 0); X[4] = cc.addc_cc(X[6], 0); X[5] = cc.addc_cc(X[7], 0)
            X[6] = cc.addc(0, 0); X[7] = 0
        m = (Y[0] * inv) & M32
        # X += p_odd * m (no carry-out), Y += p_even * m (carry into X[7])
        X[0] = cc.mad_lo_cc(po[0], m, X[0]); X[1] = cc.madc_hi_cc(po[0], m, X[1])
        X[2] = cc.madc_lo_cc(po[1], m, X[2]); X[3] = cc.madc_hi_cc(po[1], m, X[3])
        X[4] = cc.madc_lo_cc(po[2], m, X[4]); X[5] = cc.madc_hi_cc(po[2], m, X[5])
        X[6] = cc.madc_lo_cc(po[3], m, X[6]); X[7] = cc.madc_hi(po[3], m, X[7])
        Y[0] = cc.mad_lo_cc(pe[0], m, Y[0]); Y[1] = cc.madc_hi_cc(pe[0], m, Y[1])
        Y[2] = cc.madc_lo_cc(pe[1], m, Y[2]); Y[3] = cc.madc_hi_cc(pe[1], m, Y[3])
        Y[4] = cc.madc_lo_cc(pe[2], m, Y[4]); Y[5] = cc.madc_hi_cc(pe[2], m, Y[5])
        Y[6] = cc.madc_lo_cc(pe[3], m, Y[6]); Y[7] = cc.madc_hi_cc(pe[3], m, Y[7])
        X[7] = cc.addc(X[7], 0)
        assert Y[0] == 0
    # iteration 7: Y = t[1] (even role, limb 0 zero), X = t[0] (odd role): result = X + (Y >> 32)
    Y = t[1]; X = t[0]
    r = [0] * 8
    r[0] = cc.add_cc(X[0], Y[1])
    for k in range(1, 7):
        r[k] = cc.addc_cc(X[k], Y[k + 1])
    r[7] = cc.addc(X[7], 0)
    # + T_hi, then one conditional subtraction
    r[0] = cc.add_cc(r[0], T[8])
    for k in range(1, 7):
        r[k] = cc.addc_cc(r[k], T[8 + k])
    r[7] = cc.addc(r[7], T[15])
    v = val(r)
    assert v < 2 * p
    return v - p if v >= p else v


def fp2_mul_lazy(a0, a1, b0, b1, p, inv):
    T0 = val(mul_wide(a0, b0)); T1 = val(mul_wide(a1, b1))
    sa, sb = a0 + a1, b0 + b1
    assert sa < 1 << 256 and sb < 1 << 256
    T2 = val(mul_wide(sa, sb))
    C1 = T2 - T0 - T1
    assert 0 <= C1 < p << 256
    C0 = T0 - T1
    if C0 < 0:
        C0 += p << 256
    assert 0 <= C0 < p << 256
    return redc(limbs(C0, 16), p, inv), redc(limbs(C1, 16), p, inv)


if __name__ == "__main__":
    random.seed(1)
    for mod in (P, R):
        inv = (-pow(mod, -1, 1 << 32)) % (1 << 32); Rinv = pow(1 << 256, -1, mod)
        edge = [0, 1, mod - 1, mod - 2, (1 << 255) % mod, (1 << 224) - 1]
        for t in range(3000):
            pick = lambda: random.choice(edge) if t < 400 else random.randrange(mod)
            a0, a1, b0, b1 = pick(), pick(), pick(), pick()
            c0, c1 = fp2_mul_lazy(a0, a1, b0, b1, mod, inv)
            assert c0 == (a0 * b0 - a1 * b1) * Rinv % mod and c1 == (a0 * b1 + a1 * b0) * Rinv % mod
            x = random.randrange(mod << 256) if t >= 400 else random.choice([0, (mod << 256) - 1, (1 << 256) - 1, 1 << 256])
            assert redc(limbs(x, 16), mod, inv) == x * Rinv % mod
            # unreduced operands (< 2^256) of the wide product
            u, v = random.randrange(1 << 256), random.randrange(1 << 256)
            mul_wide(u, v)
            mul_wide((1 << 256) - 1, (1 << 256) - 1)
    print("ok")
