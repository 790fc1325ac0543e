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
