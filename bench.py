#!/usr/bin/env python
"""Headline benchmark: Groth16 proofs/s for the synthetic 2^20-constraint BN254 circuit (BASELINE.json
configs[2]), fake_setup zkey, random full-width witness, fixed masks.

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA prover, one rank per GPU
  python bench.py --impl reference ...                   the reference's CPU prover restated (oracle/), host cores

One JSON line on stdout (rank 0).  A "step" is one full proof: witness -> buildABC -> quotient -> 5 MSMs ->
(pi_a, pi_b, pi_c).  `value`: witness already resident in HBM; `e2e`: witness in pinned host memory, H2D copy
and the D2H read of the proof inside the timed region, through the reference-facing call (g16_prove).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))

METRIC = "groth16_proofs_per_sec_2^20_bn254"     # headline (BASELINE.json); other --log-n values rename it


def metric_name(args):
    return "groth16_proofs_per_sec_2^%d_bn254" % args.log_n
UNIT = "proofs/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ fixture
def make_fixture(g, log_n: int, want_scalars=False):
    """Synthetic R1CS with 2^log_n - 2 constraints, fake-setup zkey on the GPU, witness."""
    t0 = time.time()
    r1cs, wit = g.synthetic_chain_circuit((1 << log_n) - 2, seed=3)
    tox = g.ToxicWaste(alpha=0x1234567 + (1 << 200), beta=0x89ABCDE + (1 << 201), gamma=0x13579B + (1 << 202),
                       delta=0x2468AC + (1 << 203), tau=0xFEDCBA + (1 << 204))
    zk, sc = g.fake_circuit_setup(r1cs, tox, g.SNARKJS, want_scalars=want_scalars)
    assert zk.logDomainSize == log_n
    log("fixture 2^%d built in %.1f s (nvars=%d, ncoeffs=%d)" % (log_n, time.time() - t0, zk.nvars, zk.coeffs.shape[0]))
    return zk, wit, sc


MASK_R = 0x0A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A
MASK_S = 0x1B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU prover (prover.nim:215-304) as restated in oracle/g16_oracle_cpu.cpp -- the reference
    itself is Nim + un-vendored constantine and cannot be built here.  Each step is a bounded
    sample: one full proof of the same circuit family at 2^sample_log_n constraints, all host threads; the
    value is scaled linearly in the constraint count to the 2^20 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    import g16b200 as g
    sl = min(args.sample_log_n, args.log_n)
    zk, wit, _ = make_fixture(g, sl)                       # fixture generation (GPU fake setup) is untimed
    cores = oc.ncpu()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oc.prove(zk, wit, MASK_R, MASK_S, nthreads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    per_sample = sum(times) / len(times)
    scale = float(1 << (args.log_n - sl))
    value = 1.0 / (per_sample * scale)
    sample = "full CPU proof at 2^%d constraints (%.2f s each), scaled x%d to 2^%d" % (sl, per_sample, int(scale), args.log_n)
    out = {"impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_sample * scale * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32x8-montgomery (u64x4 on CPU)",
           "data": "synthetic", "config": workload_config(args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args, world):
    return {"workload": "synthetic R1CS 2^%d constraints (chain circuit, nvars = 2^%d), fake_setup zkey (Snarkjs "
                        "flavour), random full-width witness, fixed masks r,s; full prove" % (args.log_n, args.log_n),
            "log_constraints": args.log_n, "curve": "bn254",
            "parallelism": "1 gpu" if world == 1 else "msm point-range shards x%d" % world,
            "l2_policy": "inputs larger than L2: every proof streams the resident window tables "
                         "(about %.1f GB at this size) against a 126 MB L2" % (13 * 6 * 64 * (1 << args.log_n) / 1e9)}


# ------------------------------------------------------------------------------------------------ main arm
def run_ours(args):
    import numpy as np
    import torch
    import g16b200 as g
    from g16b200 import _lib, encoding as E
    from g16b200.prover import MEM_DEVICE, MEM_HOST

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("note: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    torch.cuda.set_device(local)
    lib = _lib.load()
    _lib.check(lib.g16_set_device(local))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    zk, wit, _ = make_fixture(g, args.log_n)
    mask = g.Mask(MASK_R, MASK_S)
    nvars = zk.nvars
    w_np = np.ascontiguousarray(wit, dtype=np.uint64)
    w_pinned = torch.from_numpy(w_np.view(np.int64).copy()).pin_memory()
    w_dev = w_pinned.to("cuda")
    h2d_bytes = nvars * 32
    d2h_bytes = 256

    # `depth` proofs in flight (default 2): consecutive proofs overlap, so the latency-bound tail of one
    # (bucket-reduction levels, assembly, the all-gather) hides behind the accumulation kernels of the next.
    # Each proof in flight owns a context slot; the slots share one resident key (g16_ctx_clone).  depth 1 =
    # strictly sequential proofs, reported as `sequential`.
    depth = max(1, args.pipeline)
    if world == 1:
        ctx = g.ProverContext(zk)
        ctxs = [ctx] + [ctx.clone() for _ in range(depth - 1)]

        def make_runner(ptr, mem_kind):
            def run(steps):
                last = None
                for i in range(steps):
                    c = ctxs[i % depth]
                    if i >= depth:
                        last = c.wait()[0]
                    c.submit(ptr, mask, E.FORM_STD, mem_kind)
                for i in range(min(depth, steps)):
                    last = ctxs[(steps - min(depth, steps) + i) % depth].wait()[0]
                return last
            return run
    else:
        sp0 = g.parallel.ShardedProver(zk, rank, world, device=local)
        sps = [sp0] + [g.parallel.ShardedProver(zk, rank, world, device=local, share=sp0) for _ in range(depth - 1)]
        ctxs = [sp.ctx for sp in sps]
        ctx = ctxs[0]

        def make_runner(ptr, mem_kind):
            def run(steps):
                last = None
                for i in range(steps):
                    sp = sps[i % depth]
                    sp.partials_submit(ptr, mem_kind, mask)
                    if i >= depth - 1 and depth > 1:
                        last = sps[(i - (depth - 1)) % depth].complete(mask)
                    elif depth == 1:
                        last = sp.complete(mask)
                for j in range(max(0, steps - (depth - 1)), steps):
                    if depth > 1:
                        last = sps[j % depth].complete(mask)
                return last
            return run

    run_resident = make_runner(w_dev.data_ptr(), MEM_DEVICE)
    run_e2e = make_runner(w_pinned.data_ptr(), MEM_HOST)

    def timed(run, steps, warmup):
        run(warmup)
        barrier()
        ms = C.c_float()
        l0 = lib.g16_kernel_launch_count()
        _lib.check(lib.g16_ctx_timer_start(ctx._h))
        t0 = time.perf_counter()
        raw = run(steps)
        torch.cuda.synchronize()
        _lib.check(lib.g16_ctx_timer_stop(ctx._h, C.byref(ms)))
        wall = (time.perf_counter() - t0) * 1e3
        launches = lib.g16_kernel_launch_count() - l0
        barrier()
        dev_ms = float(ms.value)
        if dist is not None:
            t = torch.tensor([dev_ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms, wall = float(t[0]), float(t[1])
        return dev_ms, wall, launches, raw

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms, launches, raw = timed(run_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_dev_ms, e2e_wall_ms, _, raw2 = timed(run_e2e, args.steps, max(depth, args.warmup // 2))
    seq = None
    if depth > 1:                      # strictly sequential proofs on one context, for the latency figure
        save = depth
        depth = 1
        seq_dev_ms, seq_wall_ms, _, raw3 = timed(make_runner(w_dev.data_ptr(), MEM_DEVICE), args.steps, 2)
        depth = save
        assert bytes(raw3.pi_c) == bytes(raw.pi_c)
        seq = {"ms_per_proof": seq_dev_ms / args.steps, "proofs_per_s": args.steps / (seq_dev_ms * 1e-3)}
    assert bytes(raw.pi_c) == bytes(raw2.pi_c) and bytes(raw.pi_a) == bytes(raw2.pi_a)
    stats = ctx.last_stats or {}

    out = None
    if rank == 0:
        value = args.steps / (dev_ms * 1e-3)
        out = {"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "u32x8-montgomery", "data": "synthetic",
               "config": dict(workload_config(args, world), proofs_in_flight=depth),
               "wall_ms_per_step": wall_ms / args.steps, "sequential": seq,
               "e2e": {"value": args.steps / (e2e_wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": d2h_bytes, "device_ms_per_step": e2e_dev_ms / args.steps,
                       "wall_ms_per_step": e2e_wall_ms / args.steps,
                       "api": "g16_prove_submit/wait (host witness)" if world == 1 else
                              "g16_ctx_set_mask + g16_prove_partials_submit/wait + all-gather + g16_prove_finish_submit/wait"},
               "gpu_launches": int(launches), "clocks": clocks,
               "phase_ms_last_step": {k: round(v, 3) for k, v in stats.items() if k.startswith("ms_")}}

    # ------------------------------------------------------------------ micro-benchmarks + roofline (rank 0, N = 1)
    if world == 1 and not args.no_micro:
        out.update(micro_benchmarks(args, g, lib, zk, w_dev, torch))
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_baseline(args, g, zk, wit)
        except Exception as ex:      # the checker being unavailable must not kill the GPU number
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                   "sample": "failed: %r" % (ex,)}
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def micro_benchmarks(args, g, lib, zk, w_dev, torch):
    """Standalone G1 MSM 2^k and Fr NTT 2^k (BASELINE.json configs[2]) with inputs resident in HBM, plus the
    roofline of the dominant kernel (MSM bucket accumulation) and of the NTT passes."""
    from g16b200 import _lib
    res = {}
    n = zk.nvars
    log_n = args.log_n
    # integer-pipe peak, measured live: full Montgomery multiplies / s  (kind 3) and raw mad.lo rate (kind 0)
    ops, ms = C.c_double(), C.c_float()
    _lib.check(lib.g16_bench_int_pipe(3, C.byref(ops), C.byref(ms)))
    modmul_peak = ops.value
    _lib.check(lib.g16_bench_int_pipe(0, C.byref(ops), C.byref(ms)))
    madlo_peak = ops.value
    mac_peak = modmul_peak * 136.0                 # MAC32 per Montgomery multiply

    pts = torch.from_numpy(zk.pointsA1.view("int64").copy()).to("cuda")
    result = torch.zeros(32, dtype=torch.int64, device="cuda")
    acc_ms, tot_ms, pairs = C.c_float(), C.c_float(), C.c_uint64()

    def run_msm(table: bool):
        plan = C.c_void_p()
        _lib.check(lib.g16_msm_plan_create(0, n, 0, C.byref(plan)))
        _lib.check(lib.g16_msm_plan_profile(plan, 1))
        if table:
            _lib.check(lib.g16_msm_plan_build_table(plan, pts.data_ptr(), n, None))
        accs, tots = [], []
        for i in range(args.warmup + args.steps):
            if table:
                _lib.check(lib.g16_msm_dev_table(plan, w_dev.data_ptr(), 1, n, result.data_ptr(), None))
            else:
                _lib.check(lib.g16_msm_dev(plan, w_dev.data_ptr(), 1, pts.data_ptr(), n, result.data_ptr(), None))
            _lib.check(lib.g16_msm_plan_last_profile(plan, C.byref(acc_ms), C.byref(tot_ms), C.byref(pairs)))
            if i >= args.warmup:
                accs.append(acc_ms.value)
                tots.append(tot_ms.value)
        wb, nw, ws = C.c_int(), C.c_int(), C.c_size_t()
        _lib.check(lib.g16_msm_plan_info(plan, C.byref(wb), C.byref(nw), C.byref(ws)))
        lib.g16_msm_plan_destroy(plan)
        acc, tot = sum(accs) / len(accs), sum(tots) / len(tots)
        return {"n": n, "layout": "resident window table" if table else "plain points", "window_bits": wb.value,
                "windows": nw.value, "ms": tot, "mpts_per_s": n / tot / 1e3, "bucket_accumulate_ms": acc,
                "pairs": pairs.value}

    res["msm_g1_plain"] = run_msm(False)
    res["msm_g1"] = run_msm(True)          # the layout the resident prover context uses
    acc = res["msm_g1"]["bucket_accumulate_ms"]
    modmuls = res["msm_g1"]["pairs"] * 10.0       # XYZZ mixed add = 8M + 2S
    achieved = modmuls * 136.0 / (acc * 1e-3)
    res["roofline"] = {"kernel": "k_bucket_accumulate<Fp> (G1 MSM over the resident window table, XYZZ mixed adds)",
                       "bound": "imad", "achieved": achieved / 1e12, "peak": mac_peak / 1e12, "unit": "TMAC32/s",
                       "frac": achieved / mac_peak,
                       # DRAM bytes per launch of this kernel at 2^20 from the ncu --set full capture
                       # profiles/r1_v4_ncu_full_bucket_accumulate_g1.csv (1.847 GB read + 0.131 GB written);
                       # algorithmic gather = pairs x 64 B = 0.87 GB: the kernel is integer-pipe bound, 15 % of HBM
                       "traffic": 1.979e9 if log_n == 20 else None, "traffic_unit": "bytes per launch (ncu, 2^20)",
                       "algorithmic_bytes": res["msm_g1"]["pairs"] * 68.0,
                       "peak_source": "measured live: g16_bench_int_pipe(kind=3) x 136 MAC32 per Montgomery multiply "
                                      "(IMAD.WIDE.U32 is half rate; raw mad.lo.u32 rate %.2f T/s)" % (madlo_peak / 1e12),
                       "algorithmic_work": "pairs x 10 modmul x 136 MAC32"}
    # NTT
    nn = 1 << log_n
    x = torch.from_numpy(__import__("numpy").ascontiguousarray(
        __import__("g16b200").encoding.random_fr_std(nn, 6)).view("int64")).to("cuda")
    y = torch.empty_like(x)
    wk = torch.empty_like(x)
    _lib.check(lib.g16_ntt_prepare(log_n))
    st = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for i in range(args.warmup + args.steps):
        wk.copy_(x)
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            e0.record(st)
            _lib.check(lib.g16_ntt_fr_dev(wk.data_ptr(), y.data_ptr(), wk.data_ptr(), log_n, 0, st.cuda_stream))
            e1.record(st)
        torch.cuda.synchronize()
        if i >= args.warmup:
            times.append(e0.elapsed_time(e1))
    t = sum(times) / len(times)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    import math
    passes = 1 if log_n <= 11 else 1 + math.ceil((log_n - 11) / 9)
    alg_bytes = 64.0 * nn * passes
    ntt_modmul = nn / 2 * log_n
    res["ntt_fr"] = {"n": nn, "ms": t, "melem_per_s": nn / t / 1e3, "passes": passes}
    res["roofline_ntt"] = {"kernel": "k_ntt_pass (forward NTT 2^%d, %d passes)" % (log_n, passes), "bound": "hbm",
                           "achieved": alg_bytes / (t * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": alg_bytes / (t * 1e-3) / 1e9 / hbm_peak,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                           "traffic": None,
                           "imad_frac": ntt_modmul * 136.0 / (t * 1e-3) / mac_peak,
                           "note": "254-bit NTT is integer-pipe bound (SURVEY.md 8d); imad_frac is the binding roof"}
    return res


def cpu_baseline(args, g, zk_full=None, wit_full=None):
    """The oracle's C++ restatement of the reference CPU prover timed on this box's host cores: one proof of
    the benchmark's own 2^log_n fixture when --cpu-sample-log-n equals log_n (default), else a smaller
    instance of the same circuit family scaled linearly."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    sl = min(args.cpu_sample_log_n, args.log_n)
    if sl == args.log_n and zk_full is not None:
        zk, wit = zk_full, wit_full
    else:
        zk, wit, _ = make_fixture(g, sl)
    cores = oc.ncpu()
    t0 = time.perf_counter()
    _, _, _, phases = oc.prove(zk, wit, MASK_R, MASK_S, nthreads=cores)
    dt = time.perf_counter() - t0
    scale = float(1 << (args.log_n - sl))
    labels = ["building 'ABC'", "computing the quotient (FFTs)", "computing pi_A (G1 MSM)", "computing rho (G1 MSM)",
              "computing pi_B (G2 MSM)", "computing pi_C (2x G1 MSM)"]
    return {"value": 1.0 / (dt * scale), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "one full CPU proof at 2^%d constraints (%.2f s)%s; C++ restatement of the reference "
                      "decomposition (constantine unavailable)" %
                      (sl, dt, "" if scale == 1 else ", scaled x%d to 2^%d" % (int(scale), args.log_n)),
            "phase_seconds_sample": {k: round(v, 4) for k, v in zip(labels, phases)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 of the constraint count (headline: 20)")
    ap.add_argument("--sample-log-n", type=int, default=18,
                    help="--impl reference: constraints (log2) of the bounded CPU sample proved per step")
    ap.add_argument("--cpu-sample-log-n", type=int, default=20,
                    help="cpu_baseline of the main arm: constraints (log2) of the one CPU proof that is timed")
    ap.add_argument("--pipeline", type=int, default=2, help="proofs in flight (1 = strictly sequential proofs)")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
