"""Time every shard (rank k of G) of a sharded proof on a single GPU: python tools/shard_probe.py LOG G"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib
import bench
log_n = int(sys.argv[1]); G = int(sys.argv[2])
lib = _lib.load()
zk, wit, _ = bench.make_fixture(g, log_n)
w = torch.from_numpy(np.ascontiguousarray(wit).view(np.int64).copy()).to("cuda")
parts = torch.zeros(400, dtype=torch.uint8, device="cuda")
mask = g.Mask(bench.MASK_R, bench.MASK_S)
for shards in ([G] if G else [1, 2, 4, 8]):
    worst = 0.0
    for k in range(shards):                                  # every rank of the split, one after the other
        ctx = g.ProverContext(zk, k, shards)
        ms = C.c_float()
        for i in range(4):
            torch.cuda.synchronize()
            ctx.set_mask(mask)
            _lib.check(lib.g16_ctx_timer_start(ctx._h))
            st = ctx.prove_partials(w.data_ptr(), 1, 1, parts.data_ptr())
            _lib.check(lib.g16_ctx_timer_stop(ctx._h, C.byref(ms)))
        worst = max(worst, ms.value)
        print("shards", shards, "rank", k, "partials: dev %.2f ms" % ms.value,
              {n: round(v, 2) for n, v in st.items() if n.startswith("ms_") and v}, flush=True)
        ctx.close()
    print("shards", shards, "slowest rank %.2f ms" % worst, flush=True)
