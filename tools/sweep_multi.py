"""Multi-GPU MSM / NTT size sweep (BASELINE.json configs[4]) under torchrun, one rank per GPU:
    python -m torch.distributed.run --nproc-per-node N tools/sweep_multi.py [LO HI]
G1 and G2 MSM over the resident-table layout with the points range-sharded over the ranks ([n*k/N, n*(k+1)/N), the
reference's chunking msm.nim:107-111), the partial sums all-gathered (NCCL) and added; time = max over ranks of the
device time of the rank's own MSM (CUDA events) -- the exchange is one 128/256-byte record per rank.  The Fr NTT does
not shard below the north-star threshold: N independent transforms (replicas), aggregate elements/s.  One JSON line
per size on rank 0; up to 2^PARITY_MAX the summed result is checked against the closed form (sum s_i k_i) * G."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import g16b200 as g
from g16b200 import _lib, encoding as E
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 12
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 24
parity_max = int(os.environ.get("PARITY_MAX", "18"))
torch.cuda.set_device(local)
lib = _lib.load()
_lib.check(lib.g16_set_device(local))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ops, ms = C.c_double(), C.c_float()
_lib.check(lib.g16_bench_int_pipe(3, C.byref(ops), C.byref(ms)))
modmul_peak = ops.value


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def msm(g2, n_loc, d_sc, d_pts, reps=3):
    plan = C.c_void_p()
    _lib.check(lib.g16_msm_plan_create(g2, n_loc, 0, C.byref(plan)))
    _lib.check(lib.g16_msm_plan_profile(plan, 1))
    res = torch.zeros(64, dtype=torch.int64, device="cuda")
    _lib.check(lib.g16_msm_plan_build_table(plan, d_pts.data_ptr(), n_loc, None))
    a, t, p = C.c_float(), C.c_float(), C.c_uint64()
    best = (1e9, 0.0, 0)
    for i in range(reps + 1):
        if world > 1:
            dist.barrier()
        _lib.check(lib.g16_msm_dev_table(plan, d_sc.data_ptr(), 1, n_loc, res.data_ptr(), None))
        _lib.check(lib.g16_msm_plan_last_profile(plan, C.byref(a), C.byref(t), C.byref(p)))
        tm = allmax(t.value)
        if i and tm < best[0]:
            best = (tm, allmax(a.value), p.value)
    out = np.zeros(16 if g2 else 8, dtype=np.uint64)
    _lib.check(lib.g16_msm_result_to_affine(g2, res.data_ptr(), 1, out.ctypes.data))
    wb, nw = C.c_int(), C.c_int()
    _lib.check(lib.g16_msm_plan_info(plan, C.byref(wb), C.byref(nw), None))
    lib.g16_msm_plan_destroy(plan)
    return best, wb.value, out


for lg in range(lo, hi + 1):
    n = 1 << lg
    a0, a1 = (n * rank) // world, n if rank == world - 1 else (n * (rank + 1)) // world
    row = {"log_n": lg, "n_gpus": world}
    dl = E.random_fr_std(n, seed=5)[a0:a1]
    sc = E.random_fr_std(n, seed=4)[a0:a1]
    d_sc = torch.from_numpy(np.ascontiguousarray(sc).view(np.int64)).to("cuda")
    for g2 in (0, 1):
        pts = g.fixed_base_g2(dl) if g2 else g.fixed_base_g1(dl)
        d_pts = torch.from_numpy(pts.view(np.int64)).to("cuda")
        (t, acc, pairs), c, part = msm(g2, a1 - a0, d_sc, d_pts)
        name = "g2" if g2 else "g1"
        mm = pairs * (28.0 if g2 else 10.0)
        row[name] = {"ms": round(t, 4), "mpts_per_s": round(n / t / 1e3, 2), "c_rank0": c,
                     "accumulate_ms_max": round(acc, 4),
                     "accumulate_frac_of_modmul_peak_rank0": round(mm / (acc * 1e-3) / modmul_peak, 3)}
        if lg <= parity_max:
            import g16_oracle as o
            parts = torch.from_numpy(part.view(np.int64).copy()).to("cuda")
            if world > 1:
                allp = torch.empty((world, parts.numel()), dtype=torch.int64, device="cuda")
                dist.all_gather_into_tensor(allp.view(-1), parts)
            else:
                allp = parts.view(1, -1)
            if rank == 0:
                allp = allp.cpu().numpy().view(np.uint64)
                acc_pt = o.INF_G2 if g2 else o.INF_G1
                for k in range(world):
                    pk = (E.g2_from_array(allp[k]) if g2 else E.g1_from_array(allp[k]))[0]
                    acc_pt = (o.g2_add if g2 else o.g1_add)(acc_pt, pk)
                kk = E.fr_from_std(E.random_fr_std(n, seed=5))
                ss = E.fr_from_std(E.random_fr_std(n, seed=4))
                tot = sum(x * y for x, y in zip(kk, ss)) % o.R
                row[name]["closed_form_ok"] = bool(acc_pt == ((o.g2_mul(tot, o.GEN2)) if g2 else o.g1_mul(tot, o.GEN1)))
        del d_pts, pts
    x = torch.from_numpy(E.random_fr_std(n, 6 + rank).view(np.int64)).to("cuda")
    y = torch.empty_like(x)
    _lib.check(lib.g16_ntt_prepare(lg))
    st = torch.cuda.Stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for i in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(st)
        _lib.check(lib.g16_ntt_fr_dev(x.data_ptr(), y.data_ptr(), x.data_ptr(), lg, 0, st.cuda_stream))
        e1.record(st); torch.cuda.synchronize()
        tm = allmax(e0.elapsed_time(e1))
        if i:
            best = min(best, tm)
    row["ntt_replicas"] = {"ms": round(best, 4), "melem_per_s_aggregate": round(world * n / best / 1e3, 1),
                           "frac_of_modmul_peak": round(n / 2 * lg / (best * 1e-3) / modmul_peak, 3)}
    del x, y
    torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(row), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
