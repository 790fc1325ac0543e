"""CPU: the compiled oracle (oracle/g16_oracle_cpu.cpp, the CPU-baseline restatement) against the Python
oracle and the golden fixtures."""
import random

import numpy as np

import g16_oracle as o
import oracle_cpu as oc
from g16b200 import encoding as e
from g16b200 import files


def test_ntt_literal_recursion_matches_python_oracle(kat):
    assert e.fr_from_mont(oc.ntt(e.fr_mont(kat["ntt8_in"]))) == [int(v, 16) for v in kat["ntt8_out"]]
    rnd = random.Random(1)
    for lg in (0, 1, 2, 5, 9):
        n = 1 << lg
        xs = [rnd.randrange(o.R) for _ in range(n)]
        D = o.create_domain(n)
        assert e.fr_from_mont(oc.ntt(e.fr_mont(xs))) == o.forward_ntt(xs, D)
        assert e.fr_from_mont(oc.ntt(e.fr_mont(xs), inverse=True)) == o.inverse_ntt(xs, D)


def test_msm_matches_python_oracle():
    rnd = random.Random(2)
    for n in (0, 1, 5, 130, 300):
        ks = [rnd.randrange(o.R) for _ in range(n)]
        pts = [o.g1_mul(rnd.randrange(1, 1 << 70), o.GEN1) for _ in range(n)]
        if n >= 3:
            pts[1] = o.INF_G1
            pts[2] = pts[0]
        got = oc.msm_g1(e.fr_std(ks), e.g1_array(pts) if n else np.zeros((0, 8), np.uint64), nthreads=4)
        assert e.g1_from_array(got)[0] == o.msm_naive_g1(ks, pts)
    n = 9
    ks = [rnd.randrange(o.R) for _ in range(n)]
    q = [o.g2_mul(rnd.randrange(1, 1 << 40), o.GEN2) for _ in range(n)]
    assert e.g2_from_array(oc.msm_g2(e.fr_std(ks), e.g2_array(q)))[0] == o.msm_naive_g2(ks, q)


def test_prove_matches_python_oracle_and_golden(kat):
    for name, flav in (("snarkjs", 1), ("jensgroth", 0)):
        zk = files.parse_zkey_bytes(bytes.fromhex(kat[name]["zkey_hex"]))
        zk.flavour = flav
        w = files.parse_witness_bytes(bytes.fromhex(kat["wtns_hex"])).values
        az, bz, cz = oc.build_abc(zk.coeffs, w, 3)
        assert e.fr_from_mont(az) == [int(v, 16) for v in kat[name]["Az"]]
        assert e.fr_from_mont(cz) == [int(v, 16) for v in kat[name]["Cz"]]
        assert e.fr_from_mont(oc.quotient(az, bz, cz, flav)) == [int(v, 16) for v in kat[name]["qs"]]
        r, s = int(kat["mask"]["r"], 16), int(kat["mask"]["s"], 16)
        pa, pb, pc, phases = oc.prove(zk, w, r, s, nthreads=2)
        g = kat[name]["fixed"]
        assert e.g1_from_array(pa)[0] == (int(g["pi_a"][0], 16), int(g["pi_a"][1], 16))
        assert e.g2_from_array(pb)[0][0] == (int(g["pi_b"][0][0], 16), int(g["pi_b"][0][1], 16))
        assert e.g1_from_array(pc)[0] == (int(g["pi_c"][0], 16), int(g["pi_c"][1], 16))
        assert len(phases) == 6


def test_quotient_mid_size_matches_python_fast_path():
    rnd = random.Random(3)
    lg = 9
    n = 1 << lg
    Az = [rnd.randrange(o.R) for _ in range(n)]
    Bz = [rnd.randrange(o.R) for _ in range(n)]
    Cz = [a * b % o.R for a, b in zip(Az, Bz)]
    for flav in (1, 0):
        want = (o.compute_snarkjs_scalar_coeffs if flav else o.compute_quotient_pointwise)((Az, Bz, Cz), fast=True)
        assert e.fr_from_mont(oc.quotient(e.fr_mont(Az), e.fr_mont(Bz), e.fr_mont(Cz), flav)) == want
