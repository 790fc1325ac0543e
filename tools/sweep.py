"""MSM / NTT size sweep (BASELINE.json configs[4]): G1 and G2 MSM (resident-table and plain layouts) and the
forward Fr NTT for 2^LO .. 2^HI on one GPU, with the CPU restatement beside them at the sizes it finishes
quickly.  One JSON line per size.  python tools/sweep.py [LO HI]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib, encoding as E
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 12
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cpu_max = int(os.environ.get("SWEEP_CPU_MAX", "18"))
lib = _lib.load()
ops, ms = C.c_double(), C.c_float()
_lib.check(lib.g16_bench_int_pipe(3, C.byref(ops), C.byref(ms)))
modmul_peak = ops.value
try:
    import oracle_cpu as oc
except Exception:
    oc = None

def msm(g2, n, d_sc, d_pts, table, reps=3):
    plan = C.c_void_p()
    _lib.check(lib.g16_msm_plan_create(g2, n, 0, C.byref(plan)))
    _lib.check(lib.g16_msm_plan_profile(plan, 1))
    res = torch.zeros(64, dtype=torch.int64, device="cuda")
    if table:
        _lib.check(lib.g16_msm_plan_build_table(plan, d_pts.data_ptr(), n, None))
    a, t, p = C.c_float(), C.c_float(), C.c_uint64()
    best = (1e9, 0, 0)
    for i in range(reps + 1):
        if table:
            _lib.check(lib.g16_msm_dev_table(plan, d_sc.data_ptr(), 1, n, res.data_ptr(), None))
        else:
            _lib.check(lib.g16_msm_dev(plan, d_sc.data_ptr(), 1, d_pts.data_ptr(), n, res.data_ptr(), None))
        _lib.check(lib.g16_msm_plan_last_profile(plan, C.byref(a), C.byref(t), C.byref(p)))
        if i and t.value < best[0]:
            best = (t.value, a.value, p.value)
    out = np.zeros(16 if g2 else 8, dtype=np.uint64)
    _lib.check(lib.g16_msm_result_to_affine(g2, res.data_ptr(), 1, out.ctypes.data))
    wb, nw = C.c_int(), C.c_int()
    _lib.check(lib.g16_msm_plan_info(plan, C.byref(wb), C.byref(nw), None))
    lib.g16_msm_plan_destroy(plan)
    mm = best[2] * (28.0 if g2 else 10.0)
    return {"ms": round(best[0], 4), "mpts_per_s": round(n / best[0] / 1e3, 2), "c": wb.value,
            "accumulate_ms": round(best[1], 4), "accumulate_frac_of_modmul_peak": round(mm / (best[1] * 1e-3) / modmul_peak, 3)}, out

for lg in range(lo, hi + 1):
    n = 1 << lg
    row = {"log_n": lg}
    dl = E.random_fr_std(n, seed=5)
    sc = E.random_fr_std(n, seed=4)
    d_sc = torch.from_numpy(sc.view(np.int64)).to("cuda")
    for g2 in (0, 1):
        if g2 and lg > int(os.environ.get("SWEEP_G2_MAX", "24")):
            continue
        pts = g.fixed_base_g2(dl) if g2 else g.fixed_base_g1(dl)
        d_pts = torch.from_numpy(pts.view(np.int64)).to("cuda")
        name = "g2" if g2 else "g1"
        row[name + "_table"], r1 = msm(g2, n, d_sc, d_pts, True)
        row[name + "_plain"], r2 = msm(g2, n, d_sc, d_pts, False)
        row[name + "_layouts_agree"] = bool(np.array_equal(r1, r2))
        if oc is not None and lg <= cpu_max - (2 if g2 else 0):
            t0 = time.perf_counter()
            rc = (oc.msm_g2 if g2 else oc.msm_g1)(sc, pts)
            row[name + "_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            row[name + "_cpu_matches"] = bool(np.array_equal(rc, r1))
        del d_pts, pts
    # NTT
    x = torch.from_numpy(E.random_fr_std(n, 6).view(np.int64)).to("cuda")
    y = torch.empty_like(x); wk = torch.empty_like(x)
    _lib.check(lib.g16_ntt_prepare(lg))
    st = torch.cuda.Stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for i in range(4):
        wk.copy_(x); torch.cuda.synchronize()
        e0.record(st)
        _lib.check(lib.g16_ntt_fr_dev(wk.data_ptr(), y.data_ptr(), wk.data_ptr(), lg, 0, st.cuda_stream))
        e1.record(st); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    passes = 1 if lg <= 11 else 1 + -(-(lg - 11) // 9)
    row["ntt"] = {"ms": round(best, 4), "melem_per_s": round(n / best / 1e3, 1), "passes": passes,
                  "hbm_gbs": round(64.0 * n * passes / (best * 1e-3) / 1e9, 1),
                  "frac_of_modmul_peak": round(n / 2 * lg / (best * 1e-3) / modmul_peak, 3)}
    if oc is not None and lg <= cpu_max + 2:
        xm = x.cpu().numpy().view(np.uint64)
        t0 = time.perf_counter(); rc = oc.ntt(xm); row["ntt_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        row["ntt_cpu_matches"] = bool(np.array_equal(rc, y.cpu().numpy().view(np.uint64).reshape(-1, 4)))
    del x, y, wk
    torch.cuda.empty_cache()
    print(json.dumps(row), flush=True)
