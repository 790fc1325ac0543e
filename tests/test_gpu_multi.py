"""GPU: what round 2 added around the prover -- the in-library multi-device context (one process, N devices behind
the unchanged generateProofWithMask-shaped call), the MSM-level shard plan, loader validation of the prover points
(io.nim:228-236 / curves.nim:54-107), the one-shot (plain layout) context, the mask tag of the partial records, and
the real N-process NCCL path (skipped on a one-GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import g16_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    import g16b200
    g16b200._lib.load()
    return g16b200


def E():
    from g16b200 import encoding
    return encoding


def _fixture(g, neqs=3000, flavour=1, seed=3):
    r1cs, wit = g.synthetic_chain_circuit(neqs, seed=seed)
    zk, _ = g.fake_circuit_setup(r1cs, g.ToxicWaste(11, 22, 33, 44, 55), flavour)
    return zk, np.ascontiguousarray(wit)


def _same(a, b):
    return np.array_equal(a.pi_a, b.pi_a) and np.array_equal(a.pi_b, b.pi_b) and np.array_equal(a.pi_c, b.pi_c)


def _device_count(g):
    import ctypes as C
    n = C.c_int()
    g._lib.check(g._lib.load().g16_device_count(C.byref(n)))
    return n.value


@pytest.mark.parametrize("ndev", [1, 2, 3, 8])
def test_in_library_multi_device_context_matches_single(g, ndev, monkeypatch):
    """g16_ctx_create(zkey, 0, -N): the whole key over N shards of one process; g16_prove, g16_prove_submit/wait and
    g16_ctx_clone work unchanged and give the single-device proof.  On a box with fewer than N GPUs the shards are
    placed on the devices that exist (G16_DEVICES lists device 0 several times): the same code path -- per-shard
    witness intervals, peer copies of the 400-byte records, event-ordered finish -- on fewer devices."""
    e = E()
    have = _device_count(g)
    monkeypatch.setenv("G16_DEVICES", ",".join(str(k % have) for k in range(ndev)))
    for flavour in (1, 0):
        zk, wit = _fixture(g, 2500, flavour)
        m = g.Mask(o.Rng(7).fr(), o.Rng(8).fr())
        one = g.ProverContext(zk)
        want = one.prove(wit, m)
        want0 = one.prove(wit, g.Mask(0, 0))
        one.close()
        ctx = g.ProverContext(zk, devices=ndev)
        assert _same(ctx.prove(wit, m), want)
        assert _same(ctx.prove(wit, g.Mask(0, 0)), want0)               # generateProofWithTrivialMask
        assert _same(ctx.prove(e.fr_mont(e.fr_from_std(wit)), m, e.FORM_MONT), want)
        nbytes = ctx.last_witness_bytes()
        assert nbytes >= wit.shape[0] * 32                               # at least one shard reads all of it
        # two proofs in flight over the same resident shards
        c2 = ctx.clone()
        ctx.submit(wit.ctypes.data, m)
        c2.submit(wit.ctypes.data, g.Mask(0, 0))
        assert bytes(ctx.wait()[0].pi_c) == want.pi_c.tobytes()
        raw = c2.wait()[0]
        assert bytes(raw.pi_a) == want0.pi_a.tobytes() and bytes(raw.pi_c) == want0.pi_c.tobytes()
        with pytest.raises(g._lib.G16Error):                             # shard-level calls are not for this context
            ctx.set_mask(m)
        c2.close()
        ctx.close()


def test_env_ngpus_turns_the_plain_call_multi_device(g, monkeypatch):
    """SURVEY 5 / VERDICT: with G16_NGPUS=N the unchanged drop-in call (generate_proof_with_mask, which creates its
    own context like prover.nim:215 would) runs on N shards."""
    have = _device_count(g)
    monkeypatch.setenv("G16_NGPUS", "4")
    monkeypatch.setenv("G16_DEVICES", ",".join(str(k % have) for k in range(4)))
    zk, wit = _fixture(g, 1500)
    m = g.Mask(o.Rng(9).fr(), o.Rng(10).fr())
    from g16b200.zkey_types import Witness
    got = g.generate_proof_with_mask(1, False, zk, Witness(values=wit), m)
    monkeypatch.delenv("G16_NGPUS")
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    ctx.close()
    assert _same(got, want)


def test_shard_plan_contexts_recombine_and_read_only_their_witness_slices(g):
    """One context per shard of g16_shard_plan (what one process per GPU creates): the records recombine to the
    unsharded proof, a rank without H points copies only the witness intervals of its MSM pieces, and the garbage
    outside those intervals is never read."""
    import torch
    e = E()
    from g16b200.parallel import shard_plan
    zk, wit = _fixture(g, 6000)
    m = g.Mask(o.Rng(1).fr(), o.Rng(2).fr())
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    ctx.close()
    for G in (2, 4, 8):
        parts = torch.zeros((G, g._lib.PARTIALS_BYTES), dtype=torch.uint8, device="cuda")
        ctxs = [g.ProverContext(zk, k, G) for k in range(G)]
        saw_slice = False
        for k, c in enumerate(ctxs):
            p = shard_plan(zk.nvars, zk.npubs, zk.domainSize, k, G)
            w = wit.copy()
            if p["h_hi"] == p["h_lo"]:                      # poison what this rank must not read
                keep = np.zeros(zk.nvars, dtype=bool)
                for nm in ("a1", "b1", "c1", "b2"):
                    keep[p[nm + "_lo"]:p[nm + "_hi"]] = True
                w[~keep] = 0xDEADBEEFDEADBEEF
            c.set_mask(m)
            c.prove_partials(w.ctypes.data, e.FORM_STD, 0, parts[k].data_ptr())
            if p["h_hi"] == p["h_lo"]:
                assert c.last_witness_bytes() == int(keep.sum()) * 32      # exactly the union of its pieces
                saw_slice = saw_slice or c.last_witness_bytes() < zk.nvars * 32
            else:
                assert c.last_witness_bytes() == zk.nvars * 32
        assert saw_slice or G == 2, G
        raw = ctxs[0].prove_finish(parts.data_ptr(), G, m)
        got = ctxs[0]._proof(raw, wit, e.FORM_STD)
        assert _same(got, want), G
        for c in ctxs:
            c.close()


def test_mixed_mask_conventions_are_refused(g):
    """a partial record carries whether it was produced after g16_ctx_set_mask and a hash of (r, s); the
    finish refuses a mix of masked and plain records, records of another mask, and an announced mask serves one
    set of partial sums only."""
    import torch
    e = E()
    zk, wit = _fixture(g, 900)
    m, m2 = g.Mask(o.Rng(3).fr(), o.Rng(4).fr()), g.Mask(o.Rng(5).fr(), o.Rng(6).fr())
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    ctx.close()
    G = 2
    ctxs = [g.ProverContext(zk, k, G) for k in range(G)]
    parts = torch.zeros((G, g._lib.PARTIALS_BYTES), dtype=torch.uint8, device="cuda")
    # rank 0 masked, rank 1 plain
    ctxs[0].set_mask(m)
    ctxs[0].prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[0].data_ptr())
    ctxs[1].prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[1].data_ptr())
    with pytest.raises(g._lib.G16Error, match="disagree on the mask"):
        ctxs[0].prove_finish(parts.data_ptr(), G, m)
    # both masked, but with different masks
    ctxs[0].set_mask(m)
    ctxs[0].prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[0].data_ptr())
    ctxs[1].set_mask(m2)
    ctxs[1].prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[1].data_ptr())
    with pytest.raises(g._lib.G16Error, match="disagree on the mask"):
        ctxs[0].prove_finish(parts.data_ptr(), G, m)
    # a mask announced once does not leak into the next set of partial sums: these records are plain again
    for k in range(G):
        ctxs[k].prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[k].data_ptr())
    tags = parts.cpu().numpy()[:, 384:400].view(np.uint64)
    assert not tags.any()
    raw = ctxs[1].prove_finish(parts.data_ptr(), G, m)
    assert _same(ctxs[1]._proof(raw, wit, e.FORM_STD), want)
    # and the regular masked flow still works afterwards
    for k in range(G):
        ctxs[k].set_mask(m)
        ctxs[k].prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[k].data_ptr())
    assert parts.cpu().numpy()[:, 384:392].view(np.uint64).tolist() == [[1], [1]]
    raw = ctxs[0].prove_finish(parts.data_ptr(), G, m)
    assert _same(ctxs[0]._proof(raw, wit, e.FORM_STD), want)
    for c in ctxs:
        c.close()


def test_loader_refuses_points_off_the_curve(g, kat):
    """io.nim:228-236 loadPointG1/G2 -> curves.nim:95-107 mkG1/mkG2: a prover point that does not satisfy the curve
    equation aborts the load.  Here: one flipped byte in each prover array of the golden zkey -> g16_ctx_create
    fails with the reference's message and the offending index; the point at infinity passes; G16_ZKEY_TRUSTED skips
    the check."""
    zk = g.files.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    g.ProverContext(zk).close()                                       # the intact key loads
    for name, msg in (("pointsA1", "mkG1: not a G1 curve point"), ("pointsB1", "mkG1"), ("pointsC1", "mkG1"),
                      ("pointsH1", "mkG1"), ("pointsB2", "mkG2: not a G2 curve point")):
        good = getattr(zk, name)
        bad = np.array(good, copy=True)
        idx = bad.shape[0] - 1
        bad.reshape(bad.shape[0], -1)[idx, 1] ^= np.uint64(1 << 17)
        setattr(zk, name, bad)
        with pytest.raises(g._lib.G16Error, match=msg) as ei:
            g.ProverContext(zk)
        assert "%s[%d]" % (name, idx) in str(ei.value)
        c = g.ProverContext(zk, trusted=True)                         # the benchmark's stated option: no check
        c.close()
        inf = np.array(good, copy=True)
        inf.reshape(inf.shape[0], -1)[idx, :] = 0                     # (0, 0) = infinity is a valid point
        setattr(zk, name, inf)
        g.ProverContext(zk).close()
        setattr(zk, name, good)
    # non-canonical coordinate (x + p): same point for the arithmetic, refused by the loader
    bad = np.array(zk.pointsA1, copy=True)
    x = int.from_bytes(bad[0, :4].tobytes(), "little") + o.P
    if x < 1 << 256:
        bad[0, :4] = np.frombuffer(x.to_bytes(32, "little"), dtype="<u8")
        zk2 = g.files.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
        zk2.pointsA1 = bad
        with pytest.raises(g._lib.G16Error, match="mkG1"):
            g.ProverContext(zk2)


def test_one_shot_context_matches_resident_context(g, kat):
    """G16_ZKEY_ONE_SHOT (the drop-in generateProofWithMask creates a context per call, cli_main.nim:193-210): plain
    points instead of window tables, same proof -- on the golden circuit, on a synthetic one, single and sharded."""
    e = E()
    zkg = g.files.parse_zkey_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    wtg = g.files.parse_witness_bytes(bytes.fromhex(kat["wtns_hex"]))
    r, s = int(kat["mask"]["r"], 16), int(kat["mask"]["s"], 16)
    c = g.ProverContext(zkg, one_shot=True)
    prf = c.prove(wtg.values, g.Mask(r, s))
    c.close()
    fx = kat["snarkjs"]["fixed"]
    assert e.g1_from_array(prf.pi_a)[0] == (int(fx["pi_a"][0], 16), int(fx["pi_a"][1], 16))
    assert e.g1_from_array(prf.pi_c)[0] == (int(fx["pi_c"][0], 16), int(fx["pi_c"][1], 16))
    zk, wit = _fixture(g, 20000)
    m = g.Mask(o.Rng(11).fr(), o.Rng(12).fr())
    a = g.ProverContext(zk)
    want = a.prove(wit, m)
    a.close()
    b = g.ProverContext(zk, one_shot=True)
    assert _same(b.prove(wit, m), want)
    b.close()
    have = _device_count(g)
    os.environ["G16_DEVICES"] = ",".join(str(k % have) for k in range(3))
    try:
        c = g.ProverContext(zk, devices=3, one_shot=True)
        assert _same(c.prove(wit, m), want)
        c.close()
    finally:
        del os.environ["G16_DEVICES"]


def test_witness_values_not_below_r_are_reduced(g):
    """a standard-form witness element >= r (io.nim:141-145 fromBig reduces it) must not reach the MSM
    digit extraction unreduced: w and w + r give the same proof."""
    zk, wit = _fixture(g, 700)
    m = g.Mask(o.Rng(13).fr(), o.Rng(14).fr())
    ctx = g.ProverContext(zk)
    want = ctx.prove(wit, m)
    w2 = wit.copy()
    for i in (0, 5, zk.nvars - 1):
        v = int.from_bytes(w2[i].tobytes(), "little") + o.R
        w2[i] = np.frombuffer(v.to_bytes(32, "little"), dtype="<u8")
    w2[7] = np.frombuffer(((1 << 256) - 1).to_bytes(32, "little"), dtype="<u8")
    w3 = wit.copy()
    w3[7] = np.frombuffer((((1 << 256) - 1) % o.R).to_bytes(32, "little"), dtype="<u8")
    got = ctx.prove(w2, m)
    ref = ctx.prove(w3, m)
    ctx.close()
    assert np.array_equal(got.pi_a, ref.pi_a) and np.array_equal(got.pi_b, ref.pi_b)
    assert np.array_equal(got.pi_c, ref.pi_c)
    assert not np.array_equal(want.pi_a, got.pi_a)          # element 7 really changed the statement


WORKER = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(%(root)r, "nim-groth16_b200")); sys.path.insert(0, os.path.join(%(root)r, "oracle")); sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import g16b200 as g
from g16b200.prover import MEM_HOST
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
g._lib.check(g._lib.load().g16_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
zk, wit, _ = bench.make_fixture(g, 16)
mask = g.Mask(bench.MASK_R, bench.MASK_S)
sp = g.parallel.ShardedProver(zk, rank, world, device=local)
w = np.ascontiguousarray(wit)
outs = []
wp = torch.from_numpy(w.view(np.int64).copy()).pin_memory()
wgroup = dist.new_group()
for it in range(3):
    if it == 1:                       # the witness uploaded once in `world` slices and all-gathered over NVLink
        sp.partials_submit_host(wp, mask, group=wgroup)
    else:
        sp.partials_submit(w.ctypes.data, MEM_HOST, mask)
    raw = sp.complete(mask)
    outs.append(bytes(raw.pi_a) + bytes(raw.pi_b) + bytes(raw.pi_c))
assert outs[0] == outs[1] == outs[2]
ok = True
if rank == 0:
    ok = bench.check_proof_against_ground_truth(g, zk, wit, raw)
    open(%(out)r, "w").write(json.dumps({"ok": bool(ok), "world": world, "proof": outs[0].hex()}))
dist.barrier()
sp.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_ranks_produce_the_ground_truth_proof(g, world, tmp_path):
    """the real multi-process path -- torchrun, one rank per GPU, NCCL all-gather of the masked
    partial records, device-ordered finish -- checked on hardware against ground truth: rank 0's proof must equal the
    compiled CPU restatement of the reference prover byte for byte and pass the pairing verifier.  Needs `world`
    GPUs (skipped on the one-GPU box that runs `pytest -m gpu`; bench.py repeats the check at every N > 1)."""
    if _device_count(g) < world:
        pytest.skip("needs %d GPUs" % world)
    out = tmp_path / "result.json"
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(out)})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["ok"] and res["world"] == world
    # and the same bytes as the single-GPU prover in this process
    import bench
    zk, wit, _ = bench.make_fixture(g, 16)
    ctx = g.ProverContext(zk)
    prf = ctx.prove(wit, g.Mask(bench.MASK_R, bench.MASK_S))
    ctx.close()
    assert res["proof"] == (prf.pi_a.tobytes() + prf.pi_b.tobytes() + prf.pi_c.tobytes()).hex()


def test_tables_fall_back_to_plain_points_when_they_do_not_fit(g, monkeypatch):
    """the resident tables are 13x the key; a key whose tables exceed the device budget gets the plain
    layout instead of an allocation failure (G16_TABLE_BUDGET_MB forces the situation), same proof."""
    zk, wit = _fixture(g, 9000)
    m = g.Mask(o.Rng(21).fr(), o.Rng(22).fr())
    a = g.ProverContext(zk)
    assert a.layout()[0] is True
    want = a.prove(wit, m)
    big = a.layout()[1]
    a.close()
    monkeypatch.setenv("G16_TABLE_BUDGET_MB", "1")
    b = g.ProverContext(zk)
    assert b.layout()[0] is False
    assert _same(b.prove(wit, m), want)
    assert b.layout()[1] < big
    b.close()


def test_cuda_graph_mode_gives_the_same_proofs(g):
    """G16_GRAPH=1 (opt-in): the per-proof DAG replayed as one CUDA graph from the second proof of a context slot on --
    single context, clone, sharded contexts with masked records -- bit-identical proofs (run in a child process: the
    switch is read when a context slot is created)."""
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "nim-groth16_b200")); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch
import g16b200 as g, g16_oracle as o
from g16b200 import encoding as e
r1cs, wit = g.synthetic_chain_circuit(3000, seed=3)
zk, _ = g.fake_circuit_setup(r1cs, g.ToxicWaste(11, 22, 33, 44, 55), 1)
wit = np.ascontiguousarray(wit)
masks = [g.Mask(o.Rng(k).fr(), o.Rng(k + 50).fr()) for k in range(4)]
os.environ.pop("G16_GRAPH", None)
ref = g.ProverContext(zk)
want = [ref.prove(wit, m) for m in masks]
ref.close()
os.environ["G16_GRAPH"] = "1"
c = g.ProverContext(zk)
c2 = c.clone()
for rep in range(2):
    for ctx in (c, c2):
        for m, w in zip(masks, want):
            p = ctx.prove(wit, m)
            assert np.array_equal(p.pi_a, w.pi_a) and np.array_equal(p.pi_b, w.pi_b) and np.array_equal(p.pi_c, w.pi_c)
c.close(); c2.close()
G = 3
ctxs = [g.ProverContext(zk, k, G) for k in range(G)]
parts = torch.zeros((G, g._lib.PARTIALS_BYTES), dtype=torch.uint8, device="cuda")
for m, w in zip(masks, want):
    for k, ctx in enumerate(ctxs):
        ctx.set_mask(m)
        ctx.prove_partials(wit.ctypes.data, e.FORM_STD, 0, parts[k].data_ptr())
    raw = ctxs[0].prove_finish(parts.data_ptr(), G, m)
    assert bytes(raw.pi_c) == w.pi_c.tobytes() and bytes(raw.pi_a) == w.pi_a.tobytes() and bytes(raw.pi_b) == w.pi_b.tobytes()
print("graph ok")
''' % {"root": ROOT}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "graph ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
