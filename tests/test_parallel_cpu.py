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
    local = torch.full((400,), rank + 1, dtype=torch.uint8)
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
        assert rows[0][1:] == [1] * 399 and rows[1][1:] == [2] * 399


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


def _plan_pieces(nvars, npubs, n, g):
    from g16b200.parallel import shard_plan
    return [shard_plan(nvars, npubs, n, k, g) for k in range(g)]


def test_library_shard_plan_partitions_every_array(monkeypatch):
    """g16_shard_plan (host arithmetic of the C library, no GPU): for every world size and policy each of the five
    arrays is cut into contiguous, ordered, disjoint ranges that cover it exactly once (so the partial sums add up
    to the whole MSM, msm.nim:117-119); at least one rank owns H points (and with them buildABC + the quotient);
    one rank reproduces the whole job; the uniform policy is the reference's chunking of every array."""
    from g16b200.parallel import shard_range
    import subprocess, sys, json, os
    sizes = ((1, 0, 2), (5, 1, 8), (1000, 3, 1024), (1 << 20, 1, 1 << 20), ((1 << 22) + 17, 2, 1 << 23))
    for nvars, npubs, n in sizes:
        for g in (1, 2, 3, 4, 5, 8, 16):
            plans = _plan_pieces(nvars, npubs, n, g)
            for name, total in (("a1", nvars), ("b1", nvars), ("c1", nvars), ("b2", nvars), ("h", n)):
                at = 0
                for p in plans:
                    lo, hi = p[name + "_lo"], p[name + "_hi"]
                    if hi > lo:
                        assert lo == at, (name, g, plans)
                        at = hi
                assert at == total, (name, g, plans)
            assert sum(1 for p in plans if p["h_hi"] > p["h_lo"]) >= 1
            if g == 1:
                p = plans[0]
                assert all(p[k + "_lo"] == 0 for k in ("a1", "b1", "c1", "b2", "h"))
    # at the benchmark sizes no rank owns pieces of more than three witness arrays besides whole arrays, and the
    # ranks without H points do not need the whole witness
    for lg in (20, 22):
        for g in (2, 4, 8):
            plans = _plan_pieces(1 << lg, 1, 1 << lg, g)
            assert any(p["h_hi"] == p["h_lo"] for p in plans) or g == 2
    # the policy is read once per process: check the other two in child processes
    code = ("import sys, json; sys.path.insert(0, %r); from g16b200.parallel import shard_plan; "
            "print(json.dumps([shard_plan(1000, 3, 1024, k, 4) for k in range(4)]))"
            % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nim-groth16_b200"))
    for policy in ("uniform", "g2own"):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, G16_SHARD_POLICY=policy),
                             capture_output=True, text=True, check=True).stdout
        plans = json.loads(out)
        for name, total in (("a1", 1000), ("b1", 1000), ("c1", 1000), ("b2", 1000), ("h", 1024)):
            at = 0
            for p in plans:
                if p[name + "_hi"] > p[name + "_lo"]:
                    assert p[name + "_lo"] == at
                    at = p[name + "_hi"]
            assert at == total
        if policy == "uniform":
            for k, p in enumerate(plans):
                assert (p["a1_lo"], p["a1_hi"]) == shard_range(1000, k, 4) == (p["b2_lo"], p["b2_hi"])
                assert (p["h_lo"], p["h_hi"]) == shard_range(1024, k, 4)
        else:
            assert (plans[3]["b2_lo"], plans[3]["b2_hi"]) == (0, 1000)
            assert all(p["b2_hi"] == p["b2_lo"] for p in plans[:3])
            assert plans[3]["a1_hi"] == plans[3]["a1_lo"] and plans[3]["h_hi"] == plans[3]["h_lo"]


def test_masked_shard_records_recombine_in_the_oracle(kat):
    """The algebra of the multi-GPU path (g16_ctx_set_mask, k_shard_early, k_assemble_final_masked of prover.cu)
    restated with the oracle's group law on the reference's test circuit: every rank ships
    c1' = C_k + s*A_k + r*B1_k over the ranges of g16_shard_plan (inf for an MSM it owns no points of), and
        pi_c = sum c1'_k + sum H_k + s*alpha1 + r*beta1 + (r s)*delta1,   pi_a, pi_b from the plain sums
    reproduce generateProofWithMask (prover.nim:278-304) for 2, 3, 4 and 8 ranks."""
    from g16b200.parallel import shard_plan
    zk = o.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    w = [int(v, 16) if isinstance(v, str) else int(v) for v in kat["witness"]]
    r, s = int(kat["mask"]["r"], 16), int(kat["mask"]["s"], 16)
    inter = {}
    want = o.generate_proof_with_mask(zk, w, r, s, intermediates=inter)
    qs = inter["qs"]
    first = zk.npubs + 1
    for g in (2, 3, 4, 8):
        A = B1 = C = H = o.INF_G1
        B2 = o.INF_G2
        for k in range(g):
            p = shard_plan(zk.nvars, zk.npubs, zk.domainSize, k, g)
            a_k = o.msm_naive_g1(w[p["a1_lo"]:p["a1_hi"]], zk.pointsA1[p["a1_lo"]:p["a1_hi"]])
            b1_k = o.msm_naive_g1(w[p["b1_lo"]:p["b1_hi"]], zk.pointsB1[p["b1_lo"]:p["b1_hi"]])
            b2_k = o.msm_naive_g2(w[p["b2_lo"]:p["b2_hi"]], zk.pointsB2[p["b2_lo"]:p["b2_hi"]])
            c_lo, c_hi = max(p["c1_lo"], first), max(p["c1_hi"], first)      # C1[j - npubs - 1] multiplies witness[j]
            c_k = o.msm_naive_g1(w[c_lo:c_hi], zk.pointsC1[c_lo - first:c_hi - first])
            h_k = o.msm_naive_g1(qs[p["h_lo"]:p["h_hi"]], zk.pointsH1[p["h_lo"]:p["h_hi"]])
            c1p = o.g1_add(o.g1_add(c_k, o.g1_mul(s, a_k)), o.g1_mul(r, b1_k))   # the record's c1 field
            A, B1, B2 = o.g1_add(A, a_k), o.g1_add(B1, b1_k), o.g2_add(B2, b2_k)
            C, H = o.g1_add(C, c1p), o.g1_add(H, h_k)
        pi_a = o.g1_add(o.g1_add(zk.alpha1, o.g1_mul(r, zk.delta1)), A)
        pi_b = o.g2_add(o.g2_add(zk.beta2, o.g2_mul(s, zk.delta2)), B2)
        pi_c = o.g1_add(C, H)
        for term in (o.g1_mul(s, zk.alpha1), o.g1_mul(r, zk.beta1), o.g1_mul(r * s % o.R, zk.delta1)):
            pi_c = o.g1_add(pi_c, term)
        assert (pi_a, pi_b, pi_c) == (want.pi_a, want.pi_b, want.pi_c), g


def test_planned_ranks_are_balanced_by_the_fitted_model():
    """The planner (prover.cu plan_line / rank_cost) balances the modelled busy time: at the benchmark sizes the
    slowest rank of the default plan is within 16 % of the mean over ranks for 2, 4 and 8 GPUs, no rank owns a piece
    smaller than an eighth of an array, and the plan is the same in every process (the ranks never exchange it)."""
    import subprocess, sys, os, re, json
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nim-groth16_b200")
    code = ("import sys, json; sys.path.insert(0, %r); from g16b200.parallel import shard_plan\n"
            "out = {}\n"
            "for lg in (20, 22):\n"
            "    for g in (2, 4, 8):\n"
            "        out['%%d,%%d' %% (lg, g)] = [shard_plan(1 << lg, 1, 1 << lg, k, g) for k in range(g)]\n"
            "print(json.dumps(out))" % pkg)
    env = dict(os.environ, G16_PLAN_DEBUG="1")
    env.pop("G16_SHARD_POLICY", None)
    runs = [subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, check=True)
            for _ in range(2)]
    assert runs[0].stdout == runs[1].stdout
    plans = json.loads(runs[0].stdout)
    model = {}
    for m in re.finditer(r"plan rank (\d+) of (\d+): model ([0-9.]+) ms", runs[0].stderr):
        model.setdefault(int(m.group(2)), []).append(float(m.group(3)))
    # the debug lines of the two sizes follow each other per world size: 2^20 first
    for g in (2, 4, 8):
        assert len(model[g]) == 2 * g
        for costs in (model[g][:g], model[g][g:]):
            assert max(costs) <= 1.16 * sum(costs) / g, (g, costs)
    for key, ranks in plans.items():
        lg = int(key.split(",")[0])
        for p in ranks:
            for nm in ("a1", "b1", "c1", "b2"):
                size = p[nm + "_hi"] - p[nm + "_lo"]
                assert size == 0 or size > (1 << lg) // 8, (key, p)


def test_cost_model_reproduces_the_measured_shard_shapes():
    """rank_cost (prover.cu) against the committed measurements it was fitted to (profiles/r2_v9_shape_probe.log, busy
    time of 73 shard shapes on one B200 with two proofs in flight): every shape within 10 %, rms below 0.2 ms.  Runs in
    a child process because the debug line goes to the C library's stderr."""
    import subprocess, sys, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shapes = []
    for line in open(os.path.join(root, "profiles", "r2_v9_shape_probe.log")):
        m = re.match(r"shape (.+?)\s+([-0-9.,e]+)\s+([0-9.]+) ms", line)
        if m:
            shapes.append((m.group(1).strip(), m.group(2), float(m.group(3))))
    assert len(shapes) >= 70
    code = ("import sys, os; sys.path.insert(0, %r)\n"
            "from g16b200.parallel import shard_plan\n"
            "for sh in sys.argv[1:]:\n"
            "    os.environ['G16_SHARD_SHAPE'] = sh\n"
            "    shard_plan(1 << 20, 1, 1 << 20, 1, 2)\n" % os.path.join(root, "nim-groth16_b200"))
    env = dict(os.environ, G16_PLAN_DEBUG="1")
    env.pop("G16_SHARD_POLICY", None)
    run = subprocess.run([sys.executable, "-c", code] + [s[1] for s in shapes], capture_output=True, text=True, env=env,
                         check=True)
    model = [float(v) for v in re.findall(r"shape model ([0-9.]+) ms", run.stderr)]
    assert len(model) == len(shapes)
    sq = 0.0
    for (name, _, meas), mod in zip(shapes, model):
        assert abs(mod - meas) <= 0.10 * meas, (name, meas, mod)
        sq += (mod - meas) ** 2
    assert (sq / len(shapes)) ** 0.5 < 0.2
