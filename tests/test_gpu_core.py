"""GPU parity tests proper: every call goes through the C-ABI (libg16b200.so) and is compared bit-exactly
with the oracle (oracle/g16_oracle.py) or with closed-form results on seeded inputs."""
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


def enc():
    from g16b200 import encoding
    return encoding


def test_device_selftest_and_launch_counter(g):
    lib = g._lib.load()
    before = lib.g16_kernel_launch_count()
    rc = lib.g16_selftest(1234, 256)
    assert rc == 0, lib.g16_last_error()
    assert lib.g16_kernel_launch_count() > before


# ---------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("lg", list(range(1, 13)))
def test_ntt_small_sizes_vs_literal_reference_restatement(g, lg):
    """ntt.nim:55-77 / 139-161 restated literally (recursive workers) vs the CUDA NTT, sizes 2^1..2^12."""
    E = enc()
    rnd = random.Random(100 + lg)
    n = 1 << lg
    xs = [rnd.randrange(o.R) for _ in range(n)]
    D = o.create_domain(n)
    Dg = g.create_domain(n)
    fwd = g.forward_ntt(E.fr_mont(xs), Dg)
    want = o.forward_ntt(xs, D) if lg <= 10 else o.forward_ntt_fast(xs, D)
    assert E.fr_from_mont(fwd) == want
    inv = g.inverse_ntt(E.fr_mont(xs), Dg)
    want_i = o.inverse_ntt(xs, D) if lg <= 10 else o.inverse_ntt_fast(xs, D)
    assert E.fr_from_mont(inv) == want_i


def test_ntt8_golden(g, kat):
    E = enc()
    out = g.forward_ntt(E.fr_mont(kat["ntt8_in"]), g.create_domain(8))
    assert E.fr_from_mont(out) == [int(v, 16) for v in kat["ntt8_out"]]


@pytest.mark.parametrize("lg", [13, 14, 16])
def test_ntt_multi_pass_sizes(g, lg):
    E = enc()
    rnd = random.Random(lg)
    n = 1 << lg
    xs = [rnd.randrange(o.R) for _ in range(n)]
    D = o.create_domain(n)
    fwd = g.forward_ntt(E.fr_mont(xs), g.create_domain(n))
    assert E.fr_from_mont(fwd) == o.forward_ntt_fast(xs, D)
    back = g.inverse_ntt(fwd, g.create_domain(n))
    assert E.fr_from_mont(back) == xs


@pytest.mark.parametrize("lg", [20, 22])
def test_ntt_full_size_properties(g, lg):
    """Full benchmark sizes: inverse(forward(x)) == x, linearity, and a spot check of single outputs
    against the definition sum_i x_i w^(i k)."""
    E = enc()
    n = 1 << lg
    x = E.random_fr_std(n, seed=6)          # any value < r is a valid Montgomery residue
    D = g.create_domain(n)
    fx = g.forward_ntt(x, D)
    assert np.array_equal(g.inverse_ntt(fx, D), x)
    # forward(delta_1) = powers of omega; forward(c * delta_0) = constant vector
    d = np.zeros((n, 4), dtype=np.uint64)
    d[1] = E.fr_mont([1])[0]
    fd = g.forward_ntt(d, D)
    w = o.create_domain(n).domainGen
    for k in (0, 1, 2, 12345, n - 1):
        assert E.fr_from_mont(fd[k:k + 1])[0] == pow(w, k, o.R)


def test_ntt_rejects_bad_input(g):
    with pytest.raises(g._lib.G16Error):
        g.forward_ntt(np.zeros((3, 4), np.uint64), g.create_domain(4))
    with pytest.raises(g._lib.G16Error):
        g.create_domain(6)


# ---------------------------------------------------------------------------------- MSM
def _points_g1(ks):
    return g_fixed(ks, False)


def g_fixed(ks, g2):
    import g16b200
    E = enc()
    arr = E.fr_std(ks)
    return g16b200.fixed_base_g2(arr) if g2 else g16b200.fixed_base_g1(arr)


def test_fixed_base_matches_oracle_scalar_mul(g):
    E = enc()
    rnd = random.Random(7)
    ks = [0, 1, 2, o.R - 1, 255, 256, 1 << 128] + [rnd.randrange(o.R) for _ in range(9)]
    p1 = g.fixed_base_g1(E.fr_std(ks))
    assert E.g1_from_array(p1) == [o.g1_mul(k, o.GEN1) for k in ks]
    ks2 = ks[:9]
    p2 = g.fixed_base_g2(E.fr_std(ks2))
    assert E.g2_from_array(p2) == [o.g2_mul(k, o.GEN2) for k in ks2]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 127, 128, 300])
def test_msm_g1_vs_oracle_naive(g, n):
    """msm.nim:89-124 vs msmNaiveG1 (msm.nim:162-178) restated in the oracle."""
    E = enc()
    rnd = random.Random(n)
    ks = [rnd.randrange(1, o.R) for _ in range(n)]
    pts = [o.g1_mul(rnd.randrange(1, 1 << 80), o.GEN1) for _ in range(n)]
    want = o.msm_multithreaded_g1(8, ks, pts)
    got_m = g.msm_multi_threaded_g1(8, E.fr_mont(ks), E.g1_array(pts) if n else np.zeros((0, 8), np.uint64))
    got_s = g.msm_multi_threaded_g1(8, E.fr_std(ks), E.g1_array(pts) if n else np.zeros((0, 8), np.uint64),
                                    form=E.FORM_STD)
    assert E.g1_from_array(got_m)[0] == want
    assert E.g1_from_array(got_s)[0] == want


@pytest.mark.parametrize("n", [0, 1, 5, 64, 130])
def test_msm_g2_vs_oracle_naive(g, n):
    E = enc()
    rnd = random.Random(50 + n)
    ks = [rnd.randrange(1, o.R) for _ in range(n)]
    dl = [rnd.randrange(1, 1 << 60) for _ in range(n)]
    pts_arr = g.fixed_base_g2(E.fr_std(dl)) if n else np.zeros((0, 16), np.uint64)
    want = o.g2_mul(sum(k * d for k, d in zip(ks, dl)) % o.R, o.GEN2)
    got = g.msm_multi_threaded_g2(0, E.fr_mont(ks), pts_arr)
    assert E.g2_from_array(got)[0] == want
    if 0 < n <= 5:
        assert want == o.msm_naive_g2(ks, E.g2_from_array(pts_arr))


def test_msm_edge_cases(g):
    """infinity points, zero scalars, repeated points (P + P), cancelling points (P - P), scalar r-1,
    all-equal scalars (one bucket per window), result = infinity."""
    E = enc()
    rnd = random.Random(9)
    P1 = o.g1_mul(12345, o.GEN1)
    P2 = o.g1_mul(67890, o.GEN1)
    cases = [
        ([5, 7, 0, 3], [P1, o.INF_G1, P2, o.INF_G1]),
        ([1, 1], [P1, P1]),
        ([1, 1], [P1, o.g1_neg(P1)]),
        ([o.R - 1, 1], [P1, P1]),
        ([3] * 40, [P1] * 40),
        ([3] * 33, [o.g1_mul(i + 1, o.GEN1) for i in range(33)]),
        ([0, 0, 0], [P1, P2, P1]),
        ([2 ** 253, 2 ** 16 - 1, 2 ** 15, 2 ** 15 + 1], [P1, P2, P1, P2]),
    ]
    for ks, pts in cases:
        want = o.msm_naive_g1(ks, pts)
        got = g.msm_g1(E.fr_mont(ks), E.g1_array(pts))
        assert E.g1_from_array(got)[0] == want, (ks[:3], "...")
    Q = o.g2_mul(777, o.GEN2)
    for ks, pts in [([1, 1], [Q, Q]), ([1, 1], [Q, o.g2_neg(Q)]), ([4, 0], [o.INF_G2, Q])]:
        got = g.msm_g2(E.fr_mont(ks), E.g2_array(pts))
        assert E.g2_from_array(got)[0] == o.msm_naive_g2(ks, pts)


def test_msm_length_mismatch_raises(g):
    with pytest.raises(g._lib.G16Error, match="incompatible sequence lengths"):
        g.msm_g1(np.zeros((2, 4), np.uint64), np.zeros((3, 8), np.uint64))


@pytest.mark.parametrize("lg,g2", [(12, False), (16, False), (20, False), (12, True), (16, True)])
def test_msm_closed_form_large(g, lg, g2):
    """SURVEY.md 8d MSM microbench inputs: points k_i * g with known k_i, uniform scalars; expected
    result (sum s_i k_i mod r) * g.  Also a skewed scalar distribution (zeros / ones / small values)."""
    E = enc()
    n = 1 << lg
    dl = E.random_fr_std(n, seed=5)
    pts = g.fixed_base_g2(dl) if g2 else g.fixed_base_g1(dl)
    dl_i = E.fr_from_std(dl)
    for dist in ("uniform", "skewed"):
        sc = E.random_fr_std(n, seed=4)
        if dist == "skewed":
            rng = np.random.Generator(np.random.PCG64(8))
            cls = rng.integers(0, 5, size=n)
            sc[cls <= 1] = 0                                  # 40 % zeros
            sc[cls == 2] = np.array([1, 0, 0, 0], np.uint64)  # 20 % ones
            small = cls == 3
            sc[small, 1:] = 0
            sc[small, 0] &= np.uint64(0xFFFF)                 # 20 % below 2^16
        sc_i = E.fr_from_std(sc)
        tot = sum(a * b for a, b in zip(sc_i, dl_i)) % o.R
        if g2:
            got = g.msm_multi_threaded_g2(0, sc, pts, form=E.FORM_STD)
            assert E.g2_from_array(got)[0] == o.g2_mul(tot, o.GEN2)
        else:
            got = g.msm_multi_threaded_g1(0, sc, pts, form=E.FORM_STD)
            assert E.g1_from_array(got)[0] == o.g1_mul(tot, o.GEN1)


@pytest.mark.parametrize("lg,g2", [(10, False), (16, False), (10, True), (14, True)])
def test_msm_resident_table_layout(g, lg, g2):
    """The resident-key layout (window tables 2^(cw) P_i, one bucket set) against the closed form and the
    plain layout, uniform and skewed scalars, through g16_msm_plan_build_table / g16_msm_dev_table."""
    import ctypes as C
    import torch
    E = enc()
    lib = g._lib.load()
    n = 1 << lg
    dl = E.random_fr_std(n, seed=15)
    pts = g.fixed_base_g2(dl) if g2 else g.fixed_base_g1(dl)
    pts[3] = 0                                             # a point at infinity inside the table
    dl_i = E.fr_from_std(dl)
    dl_i[3] = 0
    d_pts = torch.from_numpy(pts.view(np.int64).copy()).to("cuda")
    plan = C.c_void_p()
    g._lib.check(lib.g16_msm_plan_create(1 if g2 else 0, n, 0, C.byref(plan)))
    g._lib.check(lib.g16_msm_plan_build_table(plan, d_pts.data_ptr(), n, None))
    res = torch.zeros(64, dtype=torch.int64, device="cuda")
    for dist in ("uniform", "skewed"):
        sc = E.random_fr_std(n, seed=14)
        if dist == "skewed":
            sc[: n // 2] = np.array([1, 0, 0, 0], np.uint64)      # half of the scalars equal 1: one giant bucket
            sc[n // 2: n // 2 + n // 4] = 0
        d_sc = torch.from_numpy(sc.view(np.int64).copy()).to("cuda")
        g._lib.check(lib.g16_msm_dev_table(plan, d_sc.data_ptr(), 1, n, res.data_ptr(), None))
        out = np.zeros(16 if g2 else 8, dtype=np.uint64)
        g._lib.check(lib.g16_msm_result_to_affine(1 if g2 else 0, res.data_ptr(), 1, out.ctypes.data))
        tot = sum(a * b for a, b in zip(E.fr_from_std(sc), dl_i)) % o.R
        if g2:
            assert E.g2_from_array(out)[0] == o.g2_mul(tot, o.GEN2)
        else:
            assert E.g1_from_array(out)[0] == o.g1_mul(tot, o.GEN1)
        plain = (g.msm_multi_threaded_g2 if g2 else g.msm_multi_threaded_g1)(0, sc, pts, form=E.FORM_STD)
        assert np.array_equal(plain, out)
    lib.g16_msm_plan_destroy(plan)


def test_msm_batched_affine_tree_mode_matches():
    """The experimental batched-affine bucket accumulation (msm_tree.cuh, G16_MSM_TREE=5: read once per process)
    must give the same bit-exact answers: rerun the MSM and prover parity tests in a child process in that mode."""
    import os
    import subprocess
    import sys
    if os.environ.get("G16_MSM_TREE"):
        pytest.skip("already running in tree mode")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exp_lib = os.path.join(root, "nim-groth16_b200", "libg16b200_exp.so")
    if not os.path.exists(exp_lib):
        pytest.skip("the experiments are not in the default library: build `make -C nim-groth16_b200 EXPERIMENTS=1`")
    env = dict(os.environ, G16_MSM_TREE="5", G16B200_LIB=exp_lib)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k",
                        "msm_edge_cases or msm_resident_table_layout or msm_g2_vs_oracle_naive or reference_circuit_golden or "
                        "synthetic_circuit_closed_form or sharded_contexts_recombine",
                        os.path.join(root, "tests", "test_gpu_core.py"), os.path.join(root, "tests", "test_gpu_prover.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
