"""CPU: the host-side file readers/writers against the oracle's byte-level restatement of the formats."""
import numpy as np

import g16_oracle as o
from g16b200 import files
from g16b200.encoding import COEFF_DTYPE, fr_from_std, g1_from_array, g2_from_array
from g16b200.fake_setup import r1cs_to_coeffs, synthetic_chain_circuit


def test_zkey_parse_matches_oracle(kat):
    raw = bytes.fromhex(kat["snarkjs"]["zkey_hex"])
    zk = files.parse_zkey_bytes(raw)
    ref = o.parse_zkey_bytes(raw)
    assert (zk.nvars, zk.npubs, zk.domainSize, zk.logDomainSize) == (8, 2, 8, 3)
    assert g1_from_array(zk.pointsA1) == ref.pointsA1
    assert g1_from_array(zk.pointsH1) == ref.pointsH1
    assert g1_from_array(zk.pointsC1) == ref.pointsC1
    assert g2_from_array(zk.pointsB2) == ref.pointsB2
    assert g1_from_array(zk.alpha1)[0] == ref.alpha1 and g2_from_array(zk.delta2)[0] == ref.delta2
    assert [(int(c["m"]), int(c["row"]), int(c["col"])) for c in zk.coeffs] == \
        [(c.matrix, c.row, c.col) for c in ref.coeffs]
    assert files.write_zkey_bytes(zk) == raw                       # writer is byte-exact


def test_wtns_and_r1cs_round_trip(kat):
    w = files.parse_witness_bytes(bytes.fromhex(kat["wtns_hex"]))
    assert fr_from_std(w.values) == o.REFERENCE_TEST_WITNESS
    assert files.write_witness_bytes(w) == bytes.fromhex(kat["wtns_hex"])
    r = files.parse_r1cs_bytes(bytes.fromhex(kat["r1cs_hex"]))
    assert (r.nWires, r.nPubOut, r.nPubIn, r.nPrivIn, r.nConstr) == (8, 1, 1, 3, 3)
    assert files.write_r1cs_bytes(r) == bytes.fromhex(kat["r1cs_hex"])
    # r1csToCoeffs (fake_setup.nim:46-66) incl. the dummy public rows, R^2-encoded values
    co = r1cs_to_coeffs(r)
    want = o.r1cs_to_coeffs(o.reference_test_r1cs())
    assert [(int(c["m"]), int(c["row"]), int(c["col"])) for c in co] == [(c.matrix, c.row, c.col) for c in want]
    r2 = o.MONT * o.MONT % o.R
    assert [int.from_bytes(c["val"].tobytes(), "little") for c in co] == [c.coeff * r2 % o.R for c in want]


def test_bad_files_raise(kat):
    import pytest
    from g16b200._lib import G16Error
    raw = bytearray(bytes.fromhex(kat["wtns_hex"]))
    raw[0] = ord("x")
    with pytest.raises(G16Error):
        files.parse_witness_bytes(bytes(raw))
    z = bytearray(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    with pytest.raises(G16Error):
        files.parse_zkey_bytes(bytes(z[:-64]) + b"\0" * 10)


def test_synthetic_chain_circuit_is_satisfied():
    r1cs, wit = synthetic_chain_circuit(37, seed=3)
    w = fr_from_std(wit)
    assert r1cs.nWires == 39 and w[0] == 1 and w[1] == w[38]
    dot = lambda m, i: sum(int.from_bytes(r1cs.vals[m][t].tobytes(), "little") * w[int(r1cs.cols[m][t])]
                           for t in range(len(r1cs.rows[m])) if int(r1cs.rows[m][t]) == i) % o.R
    for i in range(r1cs.nConstr):
        assert dot(0, i) * dot(1, i) % o.R == dot(2, i)
