"""CPU (gloo, world_size 2): the host-side plumbing of the multi-GPU split -- shard ranges follow the
reference's chunking (msm.nim:107-111) and the partial-sum records all-gather intact and in rank order."""
import os
import socket

import pytest

import g16_oracle as o


def test_shard_ranges_match_reference_chunking():
    from g16b200.parallel import shard_range
    for n in (0, 1, 7, 128, 1000, (1 << 20) + 3):
        for g in (1, 2, 3, 4, 8):
            a = 0
            for k in range(g):
                lo, hi = shard_range(n, k, g)
                b = (n * (k + 1)) // g if k < g - 1 else n           # msm.nim:107-111
                assert (lo, hi) == (a, b)
                a = b
            assert a == n


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from g16b200.parallel import gather_partials
    local = torch.full((384,), rank + 1, dtype=torch.uint8)
    local[0] = 7 * (rank + 1)
    out = gather_partials(local)
    q.put((rank, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_partials_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        rows = res[rank]
        assert len(rows) == 2 and rows[0][0] == 7 and rows[1][0] == 14
        assert rows[0][1:] == [1] * 383 and rows[1][1:] == [2] * 383


def test_sharded_partial_sums_recombine_in_the_oracle():
    """msm.nim:117-119: the sum of per-shard MSMs equals the whole MSM (what prove_finish relies on)."""
    import random
    from g16b200.parallel import shard_range
    rnd = random.Random(1)
    n = 37
    ks = [rnd.randrange(o.R) for _ in range(n)]
    pts = [o.g1_mul(rnd.randrange(1, 1 << 40), o.GEN1) for _ in range(n)]
    whole = o.msm_naive_g1(ks, pts)
    for g in (2, 4):
        acc = o.INF_G1
        for k in range(g):
            lo, hi = shard_range(n, k, g)
            acc = o.g1_add(acc, o.msm_naive_g1(ks[lo:hi], pts[lo:hi]))
        assert acc == whole


def test_library_shard_ranges_partition_every_array():
    """g16_shard_ranges (host arithmetic of the C library, no GPU): for every world size the witness ranges and
    the H ranges are contiguous, ordered, disjoint and cover [0, nvars) / [0, n); the ranks without H points are
    the ones that skip buildABC and the quotient; below four ranks the split is the reference's chunking."""
    from g16b200.parallel import shard_range, shard_ranges
    for nvars, n in ((1, 2), (5, 8), (1000, 1024), ((1 << 20) - 2 + 2, 1 << 20), ((1 << 22) + 17, 1 << 23)):
        for g in (1, 2, 3, 4, 5, 8, 16):
            va = ha = 0
            h_ranks = 0
            for k in range(g):
                v_lo, v_hi, h_lo, h_hi = shard_ranges(nvars, n, k, g)
                assert v_lo == va and v_hi >= v_lo
                va = v_hi
                if h_hi > h_lo:
                    assert h_lo == ha
                    ha = h_hi
                    h_ranks += 1
                    assert g < 4 or k < max(1, g // 4)      # from 4 ranks up the H group is among the first ranks
            assert va == nvars and ha == n and h_ranks >= 1
            if g < 4:
                for k in range(g):
                    assert shard_ranges(nvars, n, k, g) == shard_range(nvars, k, g) + shard_range(n, k, g)


def test_masked_shard_records_recombine_in_the_oracle(kat):
    """The algebra of the multi-GPU path (g16_ctx_set_mask, k_shard_early, k_assemble_final_masked of prover.cu)
    restated with the oracle's group law on the reference's test circuit: every rank ships
    c1' = C_k + s*A_k + r*B1_k over the ranges of g16_shard_ranges, and
        pi_c = sum c1'_k + sum H_k + s*alpha1 + r*beta1 + (r s)*delta1,   pi_a, pi_b from the plain sums
    reproduce generateProofWithMask (prover.nim:278-304) for the uniform split (2 ranks) and the H group (4, 8)."""
    from g16b200.parallel import shard_ranges
    zk = o.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    w = [int(v, 16) if isinstance(v, str) else int(v) for v in kat["witness"]]
    r, s = int(kat["mask"]["r"], 16), int(kat["mask"]["s"], 16)
    inter = {}
    want = o.generate_proof_with_mask(zk, w, r, s, intermediates=inter)
    qs = inter["qs"]
    first = zk.npubs + 1
    for g in (2, 4, 8):
        A = B1 = C = H = o.INF_G1
        B2 = o.INF_G2
        for k in range(g):
            v_lo, v_hi, h_lo, h_hi = shard_ranges(zk.nvars, zk.domainSize, k, g)
            a_k = o.msm_naive_g1(w[v_lo:v_hi], zk.pointsA1[v_lo:v_hi])
            b1_k = o.msm_naive_g1(w[v_lo:v_hi], zk.pointsB1[v_lo:v_hi])
            b2_k = o.msm_naive_g2(w[v_lo:v_hi], zk.pointsB2[v_lo:v_hi])
            c_lo, c_hi = max(v_lo, first), max(v_hi, first)              # C1[j - npubs - 1] multiplies witness[j]
            c_k = o.msm_naive_g1(w[c_lo:c_hi], zk.pointsC1[c_lo - first:c_hi - first])
            h_k = o.msm_naive_g1(qs[h_lo:h_hi], zk.pointsH1[h_lo:h_hi])
            c1p = o.g1_add(o.g1_add(c_k, o.g1_mul(s, a_k)), o.g1_mul(r, b1_k))   # the record's c1 field
            A, B1, B2 = o.g1_add(A, a_k), o.g1_add(B1, b1_k), o.g2_add(B2, b2_k)
            C, H = o.g1_add(C, c1p), o.g1_add(H, h_k)
        pi_a = o.g1_add(o.g1_add(zk.alpha1, o.g1_mul(r, zk.delta1)), A)
        pi_b = o.g2_add(o.g2_add(zk.beta2, o.g2_mul(s, zk.delta2)), B2)
        pi_c = o.g1_add(C, H)
        for term in (o.g1_mul(s, zk.alpha1), o.g1_mul(r, zk.beta1), o.g1_mul(r * s % o.R, zk.delta1)):
            pi_c = o.g1_add(pi_c, term)
        assert (pi_a, pi_b, pi_c) == (want.pi_a, want.pi_b, want.pi_c), g
