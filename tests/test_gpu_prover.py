"""GPU: buildABC, both quotient flavours, fake setup and the full prover against the oracle, the golden
fixtures and closed-form toxic-waste identities."""
import random

import numpy as np
import pytest

import g16_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import g16b200
    g16b200._lib.load()
    return g16b200


def E():
    from g16b200 import encoding
    return encoding


def _toxic(kat):
    return {k: int(v, 16) for k, v in kat["toxic"].items()}


def _pt1(v):
    return (int(v[0], 16), int(v[1], 16))


def _pt2(v):
    return ((int(v[0][0], 16), int(v[0][1], 16)), (int(v[1][0], 16), int(v[1][1], 16)))


@pytest.mark.parametrize("name,flav", [("snarkjs", 1), ("jensgroth", 0)])
def test_reference_circuit_golden(g, kat, name, flav):
    """The reference's own test circuit (tests/groth16/testProver.nim:17-47): Az/Bz/Cz, qs and the proof
    for fixed toxic waste and masks, against tests/golden/kat.json."""
    e = E()
    zk = g.files.parse_zkey_bytes(bytes.fromhex(kat[name]["zkey_hex"]))
    zk.flavour = flav
    wt = g.files.parse_witness_bytes(bytes.fromhex(kat["wtns_hex"]))
    az, bz, cz = g.build_abc(zk, wt.values)
    assert e.fr_from_mont(az) == [int(v, 16) for v in kat[name]["Az"]]
    assert e.fr_from_mont(bz) == [int(v, 16) for v in kat[name]["Bz"]]
    assert e.fr_from_mont(cz) == [int(v, 16) for v in kat[name]["Cz"]]
    qs = (g.compute_snarkjs_scalar_coeffs if flav else g.compute_quotient_pointwise)(8, az, bz)
    assert e.fr_from_mont(qs) == [int(v, 16) for v in kat[name]["qs"]]
    # the individual MSMs through the fine-grained boundary
    got = g.msm_multi_threaded_g1(8, wt.values, zk.pointsA1, form=e.FORM_STD)
    assert e.g1_from_array(got)[0] == _pt1(kat[name]["fixed"]["msmA"])
    got = g.msm_multi_threaded_g2(8, wt.values, zk.pointsB2, form=e.FORM_STD)
    assert e.g2_from_array(got)[0] == _pt2(kat[name]["fixed"]["msmB2"])
    got = g.msm_multi_threaded_g1(8, qs, zk.pointsH1)
    assert e.g1_from_array(got)[0] == _pt1(kat[name]["fixed"]["msmH"])
    got = g.msm_multi_threaded_g1(8, wt.values[zk.npubs + 1:], zk.pointsC1, form=e.FORM_STD)
    assert e.g1_from_array(got)[0] == _pt1(kat[name]["fixed"]["msmC"])
    # the whole prover, trivial and fixed masks (prover.nim:308 / 215)
    ctx = g.ProverContext(zk)
    for masks, m in (("trivial", g.Mask(0, 0)), ("fixed", g.Mask(int(kat["mask"]["r"], 16), int(kat["mask"]["s"], 16)))):
        prf = g.generate_proof_with_mask(8, False, zk, wt, m, ctx=ctx)
        assert e.g1_from_array(prf.pi_a)[0] == _pt1(kat[name][masks]["pi_a"])
        assert e.g2_from_array(prf.pi_b)[0] == _pt2(kat[name][masks]["pi_b"])
        assert e.g1_from_array(prf.pi_c)[0] == _pt1(kat[name][masks]["pi_c"])
        assert e.fr_from_std(prf.publicIO) == [1, 2023, 1022]
        assert o.is_on_curve_g1(e.g1_from_array(prf.pi_a)[0]) and o.is_on_curve_g2(e.g2_from_array(prf.pi_b)[0])
    # Montgomery-form witness (the reference's seq[Fr]) gives the same proof
    prf2 = ctx.prove(e.fr_mont(o.REFERENCE_TEST_WITNESS), m, witness_form=e.FORM_MONT)
    assert np.array_equal(prf2.pi_c, prf.pi_c) and np.array_equal(prf2.publicIO, prf.publicIO)
    ctx.close()


@pytest.mark.parametrize("flav", [1, 0])
def test_fake_setup_matches_oracle(g, kat, flav):
    """fakeCircuitSetup (fake_setup.nim:201-326) on the GPU vs the oracle, on the reference test circuit."""
    e = E()
    r = g.files.parse_r1cs_bytes(bytes.fromhex(kat["r1cs_hex"]))
    tox = g.ToxicWaste(**_toxic(kat))
    zk, sc = g.fake_circuit_setup(r, tox, flav, want_scalars=True)
    name = "snarkjs" if flav else "jensgroth"
    assert g.files.write_zkey_bytes(zk).hex() == kat[name]["zkey_hex"]
    assert e.fr_from_std(sc.a) == [int(v, 16) for v in kat[name]["dlog_a"]]
    assert e.fr_from_std(sc.h) == [int(v, 16) for v in kat[name]["dlog_h"]]


def test_build_abc_random_sparse_matrix(g):
    """buildABC (prover.nim:56-73) with duplicate entries, empty rows and a heavy row; both record formats."""
    e = E()
    rnd = random.Random(21)
    logn, nv = 6, 50
    n = 1 << logn
    wit = [rnd.randrange(o.R) for _ in range(nv)]
    coeffs = []
    for _ in range(400):
        coeffs.append(o.Coeff(rnd.randrange(2), rnd.randrange(n // 2), rnd.randrange(nv), rnd.randrange(o.R)))
    for c in range(nv):
        coeffs.append(o.Coeff(0, 5, c, rnd.randrange(o.R)))           # heavy row
    zk_o = o.ZKey(1, nv, 1, n, logn, *[o.INF_G1] * 2, *[o.INF_G2] * 2, o.INF_G1, o.INF_G2, [], [], [], [], [], [],
                  coeffs)
    Az, Bz, Cz = o.build_abc(zk_o, wit)
    co = np.zeros(len(coeffs), dtype=e.COEFF_DTYPE)
    r2 = o.MONT * o.MONT % o.R
    co["m"], co["row"], co["col"] = [c.matrix for c in coeffs], [c.row for c in coeffs], [c.col for c in coeffs]
    co["val"] = e.ints_to_limbs(c.coeff * r2 % o.R for c in coeffs)
    zk = g.ZKey(nvars=nv, npubs=1, domainSize=n, logDomainSize=logn, flavour=1, alpha1=None, beta1=None, beta2=None,
                gamma2=None, delta1=None, delta2=None, pointsIC=None, pointsA1=None, pointsB1=None, pointsB2=None,
                pointsC1=None, pointsH1=None, coeffs=co)
    az, bz, cz = g.build_abc(zk, e.fr_std(wit))
    assert (e.fr_from_mont(az), e.fr_from_mont(bz), e.fr_from_mont(cz)) == (Az, Bz, Cz)
    az2, _, _ = g.build_abc(zk, e.fr_mont(wit), witness_form=e.FORM_MONT)
    assert np.array_equal(az, az2)
    # a matrix-C entry is fatal (prover.nim:67)
    bad = co.copy()
    bad["m"][3] = 2
    zk.coeffs = bad
    with pytest.raises(g._lib.G16Error, match="matrix C"):
        g.build_abc(zk, e.fr_std(wit))


@pytest.mark.parametrize("lg", [1, 2, 5, 12, 13])
@pytest.mark.parametrize("flav", [1, 0])
def test_quotient_vs_oracle(g, lg, flav):
    e = E()
    rnd = random.Random(lg * 2 + flav)
    n = 1 << lg
    Az = [rnd.randrange(o.R) for _ in range(n)]
    Bz = [rnd.randrange(o.R) for _ in range(n)]
    Cz = [a * b % o.R for a, b in zip(Az, Bz)]
    fast = lg > 8
    want = (o.compute_snarkjs_scalar_coeffs if flav else o.compute_quotient_pointwise)((Az, Bz, Cz), fast=fast)
    fn = g.compute_snarkjs_scalar_coeffs if flav else g.compute_quotient_pointwise
    assert e.fr_from_mont(fn(1, e.fr_mont(Az), e.fr_mont(Bz))) == want


def _closed_form_proof(g, neqs, flav, seed):
    e = E()
    r1cs, wit = g.synthetic_chain_circuit(neqs, seed=3)
    tox_o = o.ToxicWaste(*[o.Rng(seed + i).fr() or 1 for i in range(5)])
    tox = g.ToxicWaste(tox_o.alpha, tox_o.beta, tox_o.gamma, tox_o.delta, tox_o.tau)
    zk, sc = g.fake_circuit_setup(r1cs, tox, flav, want_scalars=True)
    r, s = o.Rng(seed + 10).fr(), o.Rng(seed + 11).fr()
    ctx = g.ProverContext(zk)
    prf = ctx.prove(wit, g.Mask(r, s))
    stats = ctx.last_stats
    ctx.close()
    # expected proof in the exponent; qs from the GPU quotient of the GPU-built ABC, which is
    # itself checked against the oracle at small sizes and by the verifier identity below
    az, bz, _ = g.build_abc(zk, wit)
    qs = (g.compute_snarkjs_scalar_coeffs if flav else g.compute_quotient_pointwise)(1, az, bz)
    w_i, qs_i = e.fr_from_std(wit), e.fr_from_mont(qs)
    sco = o.SetupScalars(e.fr_from_std(sc.a), e.fr_from_std(sc.b), [], e.fr_from_std(sc.ic), e.fr_from_std(sc.k),
                         e.fr_from_std(sc.h))
    cf = o.closed_form_proof_scalars(sco, tox_o, zk.npubs, w_i, qs_i, r, s)
    assert o.closed_form_check(sco, tox_o, zk.npubs, w_i, cf), "verifier equation fails in the exponent"
    ok_b = e.g2_from_array(prf.pi_b)[0] == o.g2_mul(cf["b"], o.GEN2)
    if not ok_b:      # localise: points, plain MSM, resident-table MSM or the mask term
        import random as _r
        idx = _r.Random(1).sample(range(zk.nvars), 5)
        pts_ok = all(e.g2_from_array(zk.pointsB2[j:j + 1])[0] == o.g2_mul(sco.b[j], o.GEN2) for j in idx)
        plain = e.g2_from_array(g.msm_multi_threaded_g2(0, wit, zk.pointsB2, form=e.FORM_STD))[0]
        plain_ok = plain == o.g2_mul(cf["msmB"], o.GEN2)
        ctx0 = g.ProverContext(zk)
        p0 = ctx0.prove(wit, g.Mask(0, 0))
        p1 = ctx0.prove(wit, g.Mask(r, s))
        ctx0.close()
        b0_ok = e.g2_from_array(p0.pi_b)[0] == o.g2_mul((tox_o.beta + cf["msmB"]) % o.R, o.GEN2)
        b1_ok = e.g2_from_array(p1.pi_b)[0] == o.g2_mul(cf["b"], o.GEN2)
        raise AssertionError("pi_b mismatch: B2 points ok=%s, plain MSM ok=%s, fresh ctx trivial mask ok=%s, "
                             "fresh ctx same mask ok=%s" % (pts_ok, plain_ok, b0_ok, b1_ok))
    assert e.g1_from_array(prf.pi_a)[0] == o.g1_mul(cf["a"], o.GEN1)
    assert e.g1_from_array(prf.pi_c)[0] == o.g1_mul(cf["c"], o.GEN1)
    return zk, wit, prf, stats


@pytest.mark.parametrize("neqs,flav", [(6, 1), (6, 0), (1000, 1), (1000, 0), (5000, 1)])
def test_synthetic_circuit_closed_form(g, neqs, flav):
    zk, wit, prf, _ = _closed_form_proof(g, neqs, flav, seed=40)
    if neqs <= 1000 and flav == 1:
        # and the full oracle prover on the same bytes (bit-exact on every MSM and the proof)
        e = E()
        zk_o = o.parse_zkey_bytes(g.files.write_zkey_bytes(zk))
        r, s = o.Rng(50).fr(), o.Rng(51).fr()
        if neqs <= 6:
            want = o.generate_proof_with_mask(zk_o, e.fr_from_std(wit), r, s)
            ctx = g.ProverContext(zk)
            got = ctx.prove(wit, g.Mask(r, s))
            ctx.close()
            assert e.g1_from_array(got.pi_a)[0] == want.pi_a
            assert e.g2_from_array(got.pi_b)[0] == want.pi_b
            assert e.g1_from_array(got.pi_c)[0] == want.pi_c


def test_full_prove_2_16_closed_form(g):
    """BASELINE.json configs[1]: synthetic R1CS 2^16 constraints, fake_setup zkey, random witness."""
    zk, _, _, stats = _closed_form_proof(g, (1 << 16) - 2, 1, seed=60)
    assert zk.logDomainSize == 16 and stats["kernel_launches"] > 0


def test_prover_preconditions(g, kat):
    e = E()
    zk = g.files.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    ctx = g.ProverContext(zk)
    with pytest.raises(g._lib.G16Error, match="wrong witness length"):        # prover.nim:236
        ctx.prove(np.zeros((7, 4), np.uint64), g.Mask(0, 0))
    ctx.close()
    zk.pointsH1 = zk.pointsH1[:4]
    with pytest.raises(g._lib.G16Error):                                       # prover.nim:273
        g.ProverContext(zk)


def test_sharded_contexts_recombine_on_one_gpu(g, kat):
    """The multi-GPU split exercised on one device: G shard contexts, partial records concatenated the way
    the all-gather would, g16_prove_finish == the unsharded proof (msm.nim:107-119 across devices)."""
    import torch
    e = E()
    r1cs, wit = g.synthetic_chain_circuit(700, seed=3)
    tox = g.ToxicWaste(11, 22, 33, 44, 55)
    zk, _ = g.fake_circuit_setup(r1cs, tox, 1)
    m = g.Mask(o.Rng(1).fr(), o.Rng(2).fr())
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    ctx.close()
    w = np.ascontiguousarray(wit)
    for G in (2, 3, 8):
        ctxs = [g.ProverContext(zk, k, G) for k in range(G)]
        # masked: g16_ctx_set_mask before the partial sums (every shard folds s*A_k + r*B1_k into its c1 record)
        for masked in (False, True):
            parts = torch.zeros((G, g._lib.PARTIALS_BYTES), dtype=torch.uint8, device="cuda")
            for k, c in enumerate(ctxs):
                if masked:
                    c.set_mask(m)
                c.prove_partials(w.ctypes.data, e.FORM_STD, 0, parts[k].data_ptr())
            fin = ctxs[G - 1 if masked else 0]
            raw = fin.prove_finish(parts.data_ptr(), G, m)
            got = fin._proof(raw, w, e.FORM_STD)
            assert np.array_equal(got.pi_a, want.pi_a) and np.array_equal(got.pi_b, want.pi_b)
            assert np.array_equal(got.pi_c, want.pi_c)
        # the finishing rank must be given the masks it announced
        ctxs[0].set_mask(m)
        ctxs[0].prove_partials(w.ctypes.data, e.FORM_STD, 0, parts[0].data_ptr())
        with pytest.raises(g._lib.G16Error):
            ctxs[0].prove_finish(parts.data_ptr(), G, g.Mask(m.r + 1, m.s))
        for c in ctxs:
            c.close()


def test_context_creation_is_stream_ordered_under_dirty_memory(g):
    """Regression: uploads at g16_ctx_create must be ordered before the table-building kernels on the library's
    non-blocking streams.  Recycled (dirty) device memory made a late DMA visible as a wrong G2 MSM."""
    import torch
    e = E()
    r1cs, wit = g.synthetic_chain_circuit(5000, seed=3)
    tox = g.ToxicWaste(101, 202, 303, 404, 505)
    zk, _ = g.fake_circuit_setup(r1cs, tox, 1)
    m = g.Mask(o.Rng(3).fr(), o.Rng(4).fr())
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    ctx.close()
    for rep in range(25):
        junk = torch.randint(0, 255, (rep * 1000003 % 7000000 + 10,), dtype=torch.uint8, device="cuda")
        del junk
        ctx = g.ProverContext(zk)
        got = ctx.prove(wit, m)
        ctx.close()
        assert np.array_equal(got.pi_b, want.pi_b) and np.array_equal(got.pi_a, want.pi_a)
        assert np.array_equal(got.pi_c, want.pi_c)


@pytest.mark.parametrize("flav", [1, 0])
def test_gpu_proofs_pass_the_pairing_verifier(g, flav):
    """north_star: "proofs that pass verifier.nim" -- generateProof (random masks, prover.nim:312-319) on the GPU,
    verifyProof (verifier.nim:31-52) restated with real pairings in the oracle."""
    import bn254_pairing as bp
    e = E()
    r1cs, wit = g.synthetic_chain_circuit(50, seed=9)
    zk = g.create_fake_circuit_setup(r1cs, flav)                 # random toxic waste (fake_setup.nim:330)
    ctx = g.ProverContext(zk)
    prf = g.generate_proof(4, False, zk, g.Witness(values=wit), ctx=ctx)
    prf0 = g.generate_proof_with_trivial_mask(4, False, zk, g.Witness(values=wit), ctx=ctx)
    ctx.close()
    assert not np.array_equal(prf.pi_a, prf0.pi_a)               # the masks do randomise the proof
    args = (e.g1_from_array(zk.alpha1)[0], e.g2_from_array(zk.beta2)[0], e.g2_from_array(zk.gamma2)[0],
            e.g2_from_array(zk.delta2)[0], e.g1_from_array(zk.pointsIC), e.fr_from_std(prf.publicIO))
    for p in (prf, prf0):
        assert bp.verify_proof(*args, e.g1_from_array(p.pi_a)[0], e.g2_from_array(p.pi_b)[0],
                               e.g1_from_array(p.pi_c)[0])
    # snarkjs-format export of the same proof (export_json.nim:25-80)
    txt = g.export_json.proof_json(prf)
    assert '"protocol": "groth16"' in txt and str(e.g1_from_array(prf.pi_a)[0][0]) in txt
    assert g.export_json.public_io_json(prf).count('"') == 2 * zk.npubs


def test_async_submit_wait_with_two_contexts_in_flight(g):
    """g16_prove_submit / g16_prove_wait: proofs submitted on two contexts before either is waited for are
    identical to the synchronous g16_prove results; a second submit on a busy context is refused."""
    import torch
    e = E()
    r1cs, wit = g.synthetic_chain_circuit(3000, seed=3)
    zk, _ = g.fake_circuit_setup(r1cs, g.ToxicWaste(7, 8, 9, 10, 11), 1)
    masks = [g.Mask(o.Rng(20 + i).fr(), o.Rng(30 + i).fr()) for i in range(4)]
    ref = g.ProverContext(zk)
    want = [ref.prove(wit, m) for m in masks]
    ref.close()
    first = g.ProverContext(zk)
    ctxs = [first, first.clone()]                      # g16_ctx_clone: two slots, one resident key
    w_host = torch.from_numpy(np.ascontiguousarray(wit).view(np.int64).copy()).pin_memory()
    w_dev = w_host.to("cuda")
    got = [None] * 4
    for i, m in enumerate(masks):
        c = ctxs[i % 2]
        if i >= 2:
            got[i - 2] = c.wait()[0]
        ptr, kind = (w_host.data_ptr(), 0) if i % 2 == 0 else (w_dev.data_ptr(), 1)
        c.submit(ptr, m, e.FORM_STD, kind)
    with pytest.raises(g._lib.G16Error, match="already in flight"):
        ctxs[0].submit(w_host.data_ptr(), masks[0], e.FORM_STD, 0)
    got[2] = ctxs[0].wait()[0]
    got[3] = ctxs[1].wait()[0]
    with pytest.raises(g._lib.G16Error, match="no proof in flight"):
        ctxs[0].wait()
    for raw, w in zip(got, want):
        assert bytes(raw.pi_a) == w.pi_a.tobytes() and bytes(raw.pi_b) == w.pi_b.tobytes()
        assert bytes(raw.pi_c) == w.pi_c.tobytes()
    for c in ctxs:
        c.close()


def test_sharded_prover_async_halves_single_rank(g):
    """ShardedProver.partials_submit / complete (g16_prove_partials_submit/wait + g16_prove_finish_submit) with
    world size 1 equals the plain prover."""
    import torch
    e = E()
    r1cs, wit = g.synthetic_chain_circuit(900, seed=3)
    zk, _ = g.fake_circuit_setup(r1cs, g.ToxicWaste(3, 5, 7, 11, 13), 0)
    m = g.Mask(o.Rng(5).fr(), o.Rng(6).fr())
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    ctx.close()
    sp = g.parallel.ShardedProver(zk, 0, 1, device=0)
    w = np.ascontiguousarray(wit)
    sp.partials_submit(w.ctypes.data, 0, m)
    raw = sp.complete(m)
    got = sp.prove(wit, m)
    sp.close()
    assert bytes(raw.pi_c) == want.pi_c.tobytes() and bytes(raw.pi_b) == want.pi_b.tobytes()
    assert np.array_equal(got.pi_a, want.pi_a) and np.array_equal(got.pi_c, want.pi_c)


@pytest.mark.parametrize("lg", [20, 22])
def test_full_size_proofs_pass_the_pairing_verifier(g, lg):
    """BASELINE.json's full sizes (2^20 headline, 2^22 north-star target): a size-independent property -- the
    proof of the benchmark workload (bench.make_fixture: chain circuit, GPU fake setup, full-width witness,
    fixed masks) satisfies the Groth16 verification equation of verifier.nim:31-52 (pairing restatement in the
    oracle), and the two-shard recombination (msm.nim:107-119) reproduces it byte for byte."""
    import ctypes as C
    import bench
    import bn254_pairing as bp
    e = E()
    zk, wit, _ = bench.make_fixture(g, lg)
    mask = g.Mask(bench.MASK_R, bench.MASK_S)
    ctx = g.ProverContext(zk)
    prf = ctx.prove(wit, mask)
    ctx.close()
    pub = e.fr_from_std(wit[: zk.npubs + 1])
    ok = bp.verify_proof(e.g1_from_array(zk.alpha1)[0], e.g2_from_array(zk.beta2)[0], e.g2_from_array(zk.gamma2)[0],
                         e.g2_from_array(zk.delta2)[0], e.g1_from_array(zk.pointsIC), pub,
                         e.g1_from_array(prf.pi_a)[0], e.g2_from_array(prf.pi_b)[0], e.g1_from_array(prf.pi_c)[0])
    assert ok
    if lg == 20:
        # a wrong public input must not verify
        bad = list(pub)
        bad[1] = (bad[1] + 1) % o.R
        assert not bp.verify_proof(e.g1_from_array(zk.alpha1)[0], e.g2_from_array(zk.beta2)[0],
                                   e.g2_from_array(zk.gamma2)[0], e.g2_from_array(zk.delta2)[0],
                                   e.g1_from_array(zk.pointsIC), bad, e.g1_from_array(prf.pi_a)[0],
                                   e.g2_from_array(prf.pi_b)[0], e.g1_from_array(prf.pi_c)[0])
        # the compiled CPU restatement of the reference prover (oracle/g16_oracle_cpu.cpp: its own Montgomery
        # arithmetic, chunked Pippenger, recursive NTT) on the same 2^20 inputs: bit-identical proof
        import oracle_cpu as oc
        pa, pb, pc, _ = oc.prove(zk, wit, bench.MASK_R, bench.MASK_S)
        assert np.array_equal(pa, prf.pi_a.reshape(-1)) and np.array_equal(pb, prf.pi_b.reshape(-1))
        assert np.array_equal(pc, prf.pi_c.reshape(-1))
        import torch
        w = np.ascontiguousarray(wit)
        parts = torch.zeros((2, g._lib.PARTIALS_BYTES), dtype=torch.uint8, device="cuda")
        for k in range(2):                                   # one shard context at a time: 2 x 2.6 GB of tables
            c = g.ProverContext(zk, k, 2)
            c.prove_partials(w.ctypes.data, e.FORM_STD, 0, parts[k].data_ptr())
            if k == 0:
                c.close()
        raw = c.prove_finish(parts.data_ptr(), 2, mask)
        both = c._proof(raw, w, e.FORM_STD)
        c.close()
        assert np.array_equal(both.pi_a, prf.pi_a) and np.array_equal(both.pi_b, prf.pi_b)
        assert np.array_equal(both.pi_c, prf.pi_c)


def test_full_size_intermediates_match_cpu_restatement(g):
    """north_star: "bit-exact agreement with the reference on every intermediate (H coefficients, each MSM result
    in affine form)" at the headline size 2^20: Az/Bz/Cz, qs and the five MSM results of the GPU entry points
    (g16_build_abc, g16_quotient, g16_msm_g1/g2) against the compiled CPU restatement on the same inputs."""
    import bench
    import oracle_cpu as oc
    e = E()
    zk, wit, _ = bench.make_fixture(g, 20)
    az, bz, cz = g.build_abc(zk, wit)
    az_c, bz_c, cz_c = oc.build_abc(zk.coeffs, wit, 20)
    assert np.array_equal(az, az_c) and np.array_equal(bz, bz_c) and np.array_equal(cz, cz_c)
    qs = g.compute_snarkjs_scalar_coeffs(0, az, bz)                       # prover.nim:158-181
    assert np.array_equal(qs, oc.quotient(az_c, bz_c, cz_c, 1))
    w = np.ascontiguousarray(wit)
    for pts in (zk.pointsA1, zk.pointsB1):                                # prover.nim:282,288
        assert np.array_equal(g.msm_multi_threaded_g1(0, w, pts, form=e.FORM_STD).reshape(-1), oc.msm_g1(w, pts))
    assert np.array_equal(g.msm_multi_threaded_g2(0, w, zk.pointsB2, form=e.FORM_STD).reshape(-1),
                          oc.msm_g2(w, zk.pointsB2))                      # prover.nim:294
    zs = w[zk.npubs + 1:]                                                 # prover.nim:262-264
    assert np.array_equal(g.msm_multi_threaded_g1(0, zs, zk.pointsC1, form=e.FORM_STD).reshape(-1),
                          oc.msm_g1(zs, zk.pointsC1))
    qs_std = e.fr_std(e.fr_from_mont(qs))
    assert np.array_equal(g.msm_multi_threaded_g1(0, qs, zk.pointsH1, form=e.FORM_MONT).reshape(-1),
                          oc.msm_g1(qs_std, zk.pointsH1))                 # prover.nim:301
