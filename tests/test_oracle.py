"""CPU: the oracle against the pinned vectors and its own identities."""
import random

import g16_oracle as o

# SURVEY.md Appendix C.1 / C.2 (derived independently of oracle/g16_oracle.py during the survey)
SURVEY_NTT8 = [0x344,
               0x2701a4fd3f1d3e7a309cdc72c7c8fcb5c94af009cb48e6e51461367a2f1796,
               0x2cf135e7506a45d632d270d45f1181294833fc48d823f2728,
               0x2701a4fd3f1d38dc09dff2657f0e365b7b3064279b23bdde94d81b75b0c93e,
               0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593effffffd,
               0x303d4ccde3f282f0dc4665c41c024a26ccb8b7e4521e4cd3654d1d787a4f36bb,
               0x30644e72e131a026e93ce7417adcfaf9fb0cdb0288a15dfcc0a231066dc0d8d1,
               0x303d4ccde3f282eb3e1fa8da0eb98f60726a9d586fee27aa5ecd945d75d0e863]
SURVEY_QS_SNARKJS = [0x28e8f3caa9108d0537c50b93a0484ea5c1dd002eeae0122b5f73f5761a6b8c9b,
                     0x2ef39ac2345cdb8c202f0cf6bf1a8a829f5a9c80ca35215c0e7aaeb5ca87a1a2,
                     0x14086d56217c52abcc322ef6e68d2f3671b1903bd9b84f1a023ccc2e01d7a8a8,
                     0x063c8296e783038f688986ee2430a08c9ca98a1b27b26a4efb7c0dc7c500632c,
                     0x058c9f18deb5487c65b3a2c32ddebf4018a0388232b0fda46cb6da0ba441bd1f,
                     0x13be7f75520460b73d92a29bbd84f54d4d30bea50e8dc226bda03669d239018f,
                     0x03b955206cb159172f539b38a4f3ae6eb7471dc914c5e55578a270896bf2bb95,
                     0x0206f9301fbd1f65c9a7221c8a0bfd2febf0ece260a7bfa2bd04e19b41c84abf]
SURVEY_Q_JENS = [0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593effff010,
                 0x17136539739fd53a06098765e15eff82d61bbd5a43764e8feb8f2211019f2cb6,
                 0x03f480dffd2557a578dfdd3271d443a3371408f3e0838990e694252aba670f2a,
                 0x11987d5ec68c1b7b4821f0906a96740ba31824b1ebca97d7aa9259e816303836,
                 0x10abc638c4c9639ab45a8b7065a644aefbf76db2a4a19c4f73441d402fb3be49,
                 0x0f708d184abdf7c00b5d5710502fe35e32d5557b5f0c663c84dad356de8dd121,
                 0x1bca8058ef6a4cff88774cbfa99a29bd38883870edaa7e8565f88af6f12c75c8,
                 0x0]


def test_constants_appendix_b():
    assert pow(o.GEN28, 1 << 28, o.R) == 1 and pow(o.GEN28, 1 << 27, o.R) == o.R - 1       # domain.nim:26
    assert 2 * o.ONE_HALF_FR % o.R == 1                                                     # ntt.nim:95
    assert (o.R - 1) % (1 << 28) == 0 and (o.R - 1) % (1 << 29) != 0
    assert o.MONT % o.P == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d   # io.nim:87
    assert o.MONT % o.R == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb   # io.nim:91
    assert o.inv_mod(o.MONT, o.P) == 0x2e67157159e5c639cf63e9cfb74492d9eb2022850278edf8ed84884a014afa37
    assert o.inv_mod(o.MONT, o.R) == 0x15ebf95182c5551cc8260de4aeb85d5d090ef5a9e111ec87dc5ba0056db1194e
    assert o.is_on_curve_g1(o.GEN1) and o.is_on_curve_g2(o.GEN2)
    assert o.fp2_mul((9, 1), o.TWIST_B) == (3, 0)                                           # curves.nim:75-77
    assert o.g1_mul(o.R - 1, o.GEN1) == o.g1_neg(o.GEN1)


def test_ntt8_kat(kat):
    D = o.create_domain(8)
    assert D.domainGen == 0x2b337de1c8c14f22ec9b9e2f96afef3652627366f8170a0a948dad4ac1bd5e80
    out = o.forward_ntt(list(range(101, 109)), D)
    assert out == SURVEY_NTT8 == [int(v, 16) for v in kat["ntt8_out"]]
    assert o.inverse_ntt(out, D) == list(range(101, 109))


def test_ntt_matches_naive_dft_and_fast_path():
    rnd = random.Random(5)
    for lg in range(0, 8):
        n = 1 << lg
        D = o.create_domain(n)
        xs = [rnd.randrange(o.R) for _ in range(n)]
        naive = [sum(xs[i] * pow(D.domainGen, i * k, o.R) for i in range(n)) % o.R for k in range(n)]
        assert o.forward_ntt(xs, D) == naive == o.forward_ntt_fast(xs, D)
        assert o.inverse_ntt(naive, D) == xs == o.inverse_ntt_fast(naive, D)


def test_reference_circuit_kats(kat):
    r1 = o.reference_test_r1cs()
    tox = o.ToxicWaste(**{k: int(v, 16) for k, v in kat["toxic"].items()})
    for name, fl, want in (("snarkjs", o.SNARKJS, SURVEY_QS_SNARKJS), ("jensgroth", o.JENS_GROTH, SURVEY_Q_JENS)):
        zk, sc = o.fake_circuit_setup(r1, tox, fl)
        abc = o.build_abc(zk, o.REFERENCE_TEST_WITNESS)
        assert abc[0] == [0, 7, 13, 1, 2023, 1022, 0, 0]
        assert abc[1] == [0, 11, 77, 0, 0, 0, 0, 0]
        assert abc[2] == [0, 77, 1001, 0, 0, 0, 0, 0]
        qs = o.compute_qs(zk, abc)
        assert qs == want == [int(v, 16) for v in kat[name]["qs"]]
        assert o.compute_qs(zk, abc, fast=True) == qs
        assert o.write_zkey_bytes(zk).hex() == kat[name]["zkey_hex"]


def test_closed_form_and_quotient_identity():
    rnd = random.Random(11)
    r1, wit = o.synthetic_r1cs(13, seed=9)
    tox = o.ToxicWaste(*[rnd.randrange(1, o.R) for _ in range(5)])
    for fl in (o.SNARKJS, o.JENS_GROTH):
        zk, sc = o.fake_circuit_setup(r1, tox, fl)
        r, s = rnd.randrange(o.R), rnd.randrange(o.R)
        inter = {}
        pr = o.generate_proof_with_mask(zk, wit, r, s, intermediates=inter)
        cf = o.closed_form_proof_scalars(sc, tox, zk.npubs, wit, inter["qs"], r, s)
        assert o.closed_form_check(sc, tox, zk.npubs, wit, cf)                 # verifier.nim:31-52 in the exponent
        assert pr.pi_a == o.g1_mul(cf["a"], o.GEN1)
        assert pr.pi_b == o.g2_mul(cf["b"], o.GEN2)
        assert pr.pi_c == o.g1_mul(cf["c"], o.GEN1)
        assert inter["msmH"] == o.g1_mul(cf["msmH"], o.GEN1)
    # JensGroth quotient: Q * Z == A*B - C as polynomials (top coefficient zero)
    abc = o.build_abc(zk, wit)
    q = o.compute_quotient_pointwise(abc)
    n = zk.domainSize
    assert q[-1] == 0
    D = o.create_domain(n)
    A, B, Cc = (o.inverse_ntt(v, D) for v in abc)
    x = rnd.randrange(o.R)
    ev = lambda cs: sum(c * pow(x, i, o.R) for i, c in enumerate(cs)) % o.R
    assert (ev(A) * ev(B) - ev(Cc)) % o.R == ev(q) * (pow(x, n, o.R) - 1) % o.R


def test_msm_chunking_matches_naive():
    rnd = random.Random(3)
    n = 300
    ks = [rnd.randrange(o.R) for _ in range(n)]
    pts = [o.g1_mul(rnd.randrange(1, 1 << 64), o.GEN1) for _ in range(n)]
    assert o.msm_multithreaded_g1(8, ks, pts) == o.msm_naive_g1(ks, pts)
    assert o.msm_multithreaded_g1(1, ks[:5], pts[:5]) == o.msm_naive_g1(ks[:5], pts[:5])
    assert o.msm_naive_g1([], []) == o.INF_G1


def test_file_round_trips(kat):
    zk = o.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    assert o.write_zkey_bytes(zk).hex() == kat["snarkjs"]["zkey_hex"]
    assert o.parse_wtns_bytes(bytes.fromhex(kat["wtns_hex"])) == o.REFERENCE_TEST_WITNESS
    assert zk.nvars == 8 and zk.npubs == 2 and zk.domainSize == 8 and len(zk.coeffs) == 7


# ------------------------------------------------------------------------------------------------
# Third-party BN254 (alt_bn128) vectors: values the builder did not compute.  They are the published test
# vectors of the Ethereum precompiles (EIP-196 ecAdd / ecMul cases "chfast1" of the go-ethereum / ethereum-tests
# suites) and the EIP-197 G2 generator; they pin the oracle's curve layer (curves.nim:136-154 addG1,
# :182-196 `**`) against an independent implementation.
# ------------------------------------------------------------------------------------------------
EIP196_TWO_G1 = (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
                 0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)
EIP196_ADD_CHFAST1 = (
    (0x18b18acfb4c2c30276db5411368e7185b311dd124691610c5d3b74034e093dc9,
     0x063c909c4720840cb5134cb9f59fa749755796819658d32efc0d288198f37266),
    (0x07c2b7f58a84bd6145f00c9c2bc0bb1a187f20ff2c92963a88019e7c6a014eed,
     0x06614e20c147e940f2d70da3f74c9a17df361706a4485c742bd6788478fa17d7),
    (0x2243525c5efd4b9c3d3c45ac0ca3fe4dd85e830a4ce6b65fa1eeaee202839703,
     0x301d1d33be6da8e509df21cc35964723180eed7532537db9ae5e7d48f195c915))
EIP196_MUL_CHFAST1 = (
    (0x2bd3e6d0f3b142924f5ca7b49ce5b9d54c4703d7ae5648e61d02268b1a0a9fb7,
     0x21611ce0a6af85915e2f1d70300909ce2e49dfad4a4619c8390cae66cefdb204),
    0x00000000000000000000000000000000000000000000000011138ce750fa15c2,
    (0x070a8d6a982153cae4be29d434e8faef8a47b274a053f5a4ee2a6c9c13c31e5c,
     0x031b8ce914eba3a9ffb989f9cdd5b0f01943074bf4f0f315690ec3cec6981afc))
EIP197_G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
              11559732032986387107991004021392285783925812861821192530917403151452391805634),
             (8495653923123431417604973247489272438418190587263600148770280649306958101930,
              4082367875863433681332203403145435568316851327593401208105741076214120093531))


def test_third_party_bn254_vectors_g1():
    assert o.g1_add(o.GEN1, o.GEN1) == EIP196_TWO_G1
    assert o.g1_mul(2, o.GEN1) == EIP196_TWO_G1
    assert o.msm_naive_g1([2], [o.GEN1]) == EIP196_TWO_G1
    a, b, c = EIP196_ADD_CHFAST1
    assert o.is_on_curve_g1(a) and o.is_on_curve_g1(b) and o.g1_add(a, b) == c and o.g1_add(b, a) == c
    p, k, q = EIP196_MUL_CHFAST1
    assert o.is_on_curve_g1(p) and o.g1_mul(k, p) == q
    assert o.msm_naive_g1([k, 1], [p, a]) == o.g1_add(q, a)
    assert o.g1_mul(o.R, p) == o.INF_G1 and o.g1_mul(o.R - 1, o.GEN1) == (1, o.P - 2)


def test_third_party_bn254_vectors_g2_and_pairing():
    """The EIP-197 generator of G2 (not the generator curves.nim:115-121 uses) lies on the twist and in the
    order-r subgroup under the oracle's G2 arithmetic, and pairs bilinearly with the EIP-196 points."""
    import bn254_pairing as bp
    assert o.is_on_curve_g2(EIP197_G2)
    assert o.g2_mul(o.R, EIP197_G2) == o.INF_G2
    assert o.g2_add(o.g2_mul(5, EIP197_G2), o.g2_mul(o.R - 5, EIP197_G2)) == o.INF_G2
    # the ecPairing identity of EIP-197: e(P, Q) * e(-P, Q) = 1, and e(2P, Q) = e(P, 2Q) with the published 2P
    e1 = bp.pairing(o.GEN1, EIP197_G2)
    assert not (e1 == bp.F12.one())
    assert e1 * bp.pairing(o.g1_neg(o.GEN1), EIP197_G2) == bp.F12.one()
    assert bp.pairing(EIP196_TWO_G1, EIP197_G2) == bp.pairing(o.GEN1, o.g2_mul(2, EIP197_G2)) == e1 * e1
