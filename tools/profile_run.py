"""Small fixed workload for ncu: one warm-up + one timed full proof at 2^LOG, then standalone G1 / G2 MSMs and
a forward NTT of the same size.  Usage: python tools/profile_run.py [LOG]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import g16b200 as g
from g16b200 import _lib
import bench

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = _lib.load()
zk, wit, _ = bench.make_fixture(g, log_n)
ctx = g.ProverContext(zk)
mask = g.Mask(bench.MASK_R, bench.MASK_S)
w = torch.from_numpy(np.ascontiguousarray(wit).view(np.int64).copy()).to("cuda")
for i in range(2):
    t0 = time.perf_counter()
    ctx.prove_dev(w.data_ptr(), mask)
    print("prove %d: %.2f ms" % (i, (time.perf_counter() - t0) * 1e3), ctx.last_stats, flush=True)
ctx.close()
n = zk.nvars
for g2 in (0, 1):
    pts = torch.from_numpy((zk.pointsB2 if g2 else zk.pointsA1).view("int64").copy()).to("cuda")
    plan = C.c_void_p()
    _lib.check(lib.g16_msm_plan_create(g2, n, 0, C.byref(plan)))
    _lib.check(lib.g16_msm_plan_profile(plan, 1))
    res = torch.zeros(64, dtype=torch.int64, device="cuda")
    a, t, p = C.c_float(), C.c_float(), C.c_uint64()
    for i in range(2):
        _lib.check(lib.g16_msm_dev(plan, w.data_ptr(), 1, pts.data_ptr(), n, res.data_ptr(), None))
        _lib.check(lib.g16_msm_plan_last_profile(plan, C.byref(a), C.byref(t), C.byref(p)))
        print("msm g%d plain: total %.3f ms, accumulate %.3f ms, pairs %d" % (g2 + 1, t.value, a.value, p.value), flush=True)
    t0 = time.perf_counter()
    _lib.check(lib.g16_msm_plan_build_table(plan, pts.data_ptr(), n, None))
    print("table build g%d: %.1f ms" % (g2 + 1, (time.perf_counter() - t0) * 1e3), flush=True)
    for i in range(2):
        _lib.check(lib.g16_msm_dev_table(plan, w.data_ptr(), 1, n, res.data_ptr(), None))
        _lib.check(lib.g16_msm_plan_last_profile(plan, C.byref(a), C.byref(t), C.byref(p)))
        print("msm g%d table: total %.3f ms, accumulate %.3f ms, pairs %d" % (g2 + 1, t.value, a.value, p.value), flush=True)
    lib.g16_msm_plan_destroy(plan)
x = torch.from_numpy(g.encoding.random_fr_std(1 << log_n, 6).view(np.int64)).to("cuda")
y = torch.empty_like(x)
_lib.check(lib.g16_ntt_prepare(log_n))
for i in range(2):
    wk = x.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(lib.g16_ntt_fr_dev(wk.data_ptr(), y.data_ptr(), wk.data_ptr(), log_n, 0, None))
    torch.cuda.synchronize()
    print("ntt: %.3f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
