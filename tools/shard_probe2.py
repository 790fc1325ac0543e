"""Per-rank throughput of a sharded proof measured on ONE GPU: for rank k of G the partial sums are computed in a
pipelined loop (depth proofs in flight, no exchange), which is the per-rank busy time that bounds the N-GPU rate.
    python tools/shard_probe2.py LOG G [depths...]        env: G16_C_DELTA, G16_SHARD_POLICY"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib
from g16b200.parallel import shard_plan
import bench
log_n = int(sys.argv[1]); G = int(sys.argv[2])
depths = [int(a) for a in sys.argv[3:]] or [1, 2, 3]
ranks = [int(a) for a in os.environ.get("PROBE_RANKS", "").split(",") if a] or list(range(G))
lib = _lib.load()
zk, wit, _ = bench.make_fixture(g, log_n)
w = torch.from_numpy(np.ascontiguousarray(wit).view(np.int64).copy()).to("cuda")
mask = g.Mask(bench.MASK_R, bench.MASK_S)
steps = 12
for k in ranks:
    base = g.ProverContext(zk, k, G, trusted=True)
    slots = [base] + [base.clone() for _ in range(max(depths) - 1)]
    parts = [torch.zeros(400, dtype=torch.uint8, device="cuda") for _ in slots]
    p = shard_plan(zk.nvars, zk.npubs, zk.domainSize, k, G)
    desc = " ".join("%s[%.2f,%.2f]" % (nm, p[nm + "_lo"] / (zk.domainSize if nm == "h" else zk.nvars),
                                      p[nm + "_hi"] / (zk.domainSize if nm == "h" else zk.nvars))
                    for nm in ("h", "a1", "b1", "c1", "b2") if p[nm + "_hi"] > p[nm + "_lo"])
    out = []
    for d in depths:
        def run(n):
            for i in range(n):
                c = slots[i % d]
                if i >= d:
                    _lib.check(lib.g16_prove_partials_wait(c._h, None))
                c.set_mask(mask)
                _lib.check(lib.g16_prove_partials_submit(c._h, w.data_ptr(), 1, 1, parts[i % d].data_ptr()))
            for i in range(min(d, n)):
                _lib.check(lib.g16_prove_partials_wait(slots[(n - min(d, n) + i) % d]._h, None))
        run(4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(steps)
        torch.cuda.synchronize()
        out.append("depth %d: %.2f ms" % (d, (time.perf_counter() - t0) / steps * 1e3))
    print("rank %d of %d  %-40s %s" % (k, G, desc, "  ".join(out)), flush=True)
    for c in slots:
        c.close()
