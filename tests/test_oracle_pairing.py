"""CPU: the pairing-based verifier restatement (oracle/bn254_pairing.py, verifier.nim:31-52)."""
import time

import g16_oracle as o
import bn254_pairing as bp


def test_pairing_is_bilinear_and_nondegenerate():
    e = bp.pairing(o.GEN1, o.GEN2)
    assert not (e == bp.F12.one())
    assert e.pow(o.R) == bp.F12.one()
    a, b = 5, 7
    lhs = bp.pairing(o.g1_mul(a, o.GEN1), o.g2_mul(b, o.GEN2))
    assert lhs == e.pow(a * b)
    assert bp.pairing(o.INF_G1, o.GEN2) == bp.F12.one()
    x = bp.F12([3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5, 8])
    assert x * x.inv() == bp.F12.one()


def test_verifier_accepts_golden_proofs_and_rejects_tampering(kat):
    """The golden proofs of the reference test circuit (both flavours, fixed masks) verify; a tampered
    proof or a wrong public input does not."""
    tox = {k: int(v, 16) for k, v in kat["toxic"].items()}
    p1 = lambda v: (int(v[0], 16), int(v[1], 16))
    p2 = lambda v: ((int(v[0][0], 16), int(v[0][1], 16)), (int(v[1][0], 16), int(v[1][1], 16)))
    gamma2 = o.g2_mul(tox["gamma"], o.GEN2)
    for name in ("snarkjs", "jensgroth"):
        zk = o.parse_zkey_bytes(bytes.fromhex(kat[name]["zkey_hex"]))
        pub = o.REFERENCE_TEST_WITNESS[: zk.npubs + 1]
        prf = kat[name]["fixed"]
        args = (zk.alpha1, zk.beta2, gamma2, zk.delta2, zk.pointsIC, pub)
        t0 = time.time()
        assert bp.verify_proof(*args, p1(prf["pi_a"]), p2(prf["pi_b"]), p1(prf["pi_c"]))
        assert time.time() - t0 < 120
    bad_c = o.g1_add(p1(prf["pi_c"]), o.GEN1)
    assert not bp.verify_proof(*args, p1(prf["pi_a"]), p2(prf["pi_b"]), bad_c)
    bad_pub = [1, 2024, 1022]
    assert not bp.verify_proof(zk.alpha1, zk.beta2, gamma2, zk.delta2, zk.pointsIC, bad_pub, p1(prf["pi_a"]),
                               p2(prf["pi_b"]), p1(prf["pi_c"]))
