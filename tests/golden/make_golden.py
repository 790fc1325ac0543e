"""Generates tests/golden/kat.json with the Python oracle (oracle/g16_oracle.py).

The reference cannot be run in this environment (no Nim, constantine not vendored) and holds no golden
vectors, so these fixtures are oracle-derived; the NTT-8, Az/Bz/Cz and qs values below were cross-checked
against SURVEY.md Appendix C (an independent derivation) by tests/test_oracle.py.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import g16_oracle as o  # noqa: E402

TOXIC = o.ToxicWaste(alpha=0x1111111111111111222222222222222233333333333333334444444444444444 % o.R,
                     beta=0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF % o.R,
                     gamma=0x2BADF00D2BADF00D2BADF00D2BADF00D2BADF00D2BADF00D2BADF00D2BADF00D % o.R,
                     delta=0x0FEDCBA9876543210FEDCBA9876543210FEDCBA9876543210FEDCBA987654321 % o.R,
                     tau=0x1C0FFEE01C0FFEE01C0FFEE01C0FFEE01C0FFEE01C0FFEE01C0FFEE01C0FFEE0 % o.R)
MASK_R = 0x0A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A % o.R
MASK_S = 0x1B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B % o.R

hx = lambda v: hex(v)


def pt1(p):
    return [hx(p[0]), hx(p[1])]


def pt2(p):
    return [[hx(p[0][0]), hx(p[0][1])], [hx(p[1][0]), hx(p[1][1])]]


def main():
    out = {"toxic": {k: hx(getattr(TOXIC, k)) for k in ("alpha", "beta", "gamma", "delta", "tau")},
           "mask": {"r": hx(MASK_R), "s": hx(MASK_S)}}
    D8 = o.create_domain(8)
    out["omega8"] = hx(D8.domainGen)
    out["eta16"] = hx(o.create_domain(16).domainGen)
    out["ntt8_in"] = list(range(101, 109))
    out["ntt8_out"] = [hx(v) for v in o.forward_ntt(list(range(101, 109)), D8)]
    r1 = o.reference_test_r1cs()
    wit = o.REFERENCE_TEST_WITNESS
    out["witness"] = wit
    for name, fl in (("snarkjs", o.SNARKJS), ("jensgroth", o.JENS_GROTH)):
        zk, sc = o.fake_circuit_setup(r1, TOXIC, fl)
        case = {}
        for masks, (r, s) in (("trivial", (0, 0)), ("fixed", (MASK_R, MASK_S))):
            inter = {}
            pr = o.generate_proof_with_mask(zk, wit, r, s, intermediates=inter)
            cf = o.closed_form_proof_scalars(sc, TOXIC, zk.npubs, wit, inter["qs"], r, s)
            assert o.closed_form_check(sc, TOXIC, zk.npubs, wit, cf)
            assert pr.pi_a == o.g1_mul(cf["a"], o.GEN1) and pr.pi_c == o.g1_mul(cf["c"], o.GEN1)
            assert pr.pi_b == o.g2_mul(cf["b"], o.GEN2)
            case[masks] = {"pi_a": pt1(pr.pi_a), "pi_b": pt2(pr.pi_b), "pi_c": pt1(pr.pi_c),
                           "msmA": pt1(inter["msmA"]), "msmB1": pt1(inter["msmB1"]), "msmB2": pt2(inter["msmB2"]),
                           "msmH": pt1(inter["msmH"]), "msmC": pt1(inter["msmC"])}
        case["Az"] = [hx(v) for v in inter["Az"]]
        case["Bz"] = [hx(v) for v in inter["Bz"]]
        case["Cz"] = [hx(v) for v in inter["Cz"]]
        case["qs"] = [hx(v) for v in inter["qs"]]
        case["zkey_hex"] = o.write_zkey_bytes(zk).hex()
        case["dlog_a"] = [hx(v) for v in sc.a]
        case["dlog_h"] = [hx(v) for v in sc.h]
        out[name] = case
    out["wtns_hex"] = o.write_wtns_bytes(wit).hex()
    out["r1cs_hex"] = o.write_r1cs_bytes(r1).hex()
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote kat.json")


if __name__ == "__main__":
    main()
