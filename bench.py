#!/usr/bin/env python
"""Headline benchmark: Groth16 proofs/s for the synthetic 2^20-constraint BN254 circuit (BASELINE.json
configs[2]), fake_setup zkey, random full-width witness, fixed masks.

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA prover, one rank per GPU
  python bench.py --impl reference ...                   the reference's CPU prover restated (oracle/), host cores

One JSON line on stdout (rank 0).  A "step" is one full proof: witness -> buildABC -> quotient -> 5 MSMs ->
(pi_a, pi_b, pi_c), i.e. generateProofWithMask (groth16/prover.nim:215-304).  `value`: witness already resident in
HBM; `e2e`: witness in pinned host memory, H2D copy and the D2H read of the proof inside the timed region, through
the reference-facing call.  Every line is self-verifying: the proof the timed path produced is compared outside the
timed region with the compiled CPU restatement of the reference prover on the same fixture (byte equality) and
run through the pairing verifier (`parity_checked`); a mismatch exits non-zero.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))

METRIC = "groth16_proofs_per_sec_2^20_bn254"     # headline (BASELINE.json); other --log-n values rename it
UNIT = "proofs/s"


def metric_name(args):
    return "groth16_proofs_per_sec_2^%d_bn254" % args.log_n


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


def skewed_witness(nvars: int, seed: int = 4):
    """The robustness distribution of SURVEY.md 8d: 40 % zeros, 20 % ones, 20 % below 2^16, 20 % uniform -- what real
    .wtns files look like (booleans, small counters).  Not a satisfying assignment of the chain circuit: the prover's
    work does not depend on satisfiability, so it is used for timing only."""
    import numpy as np
    from g16b200 import encoding as E
    rng = np.random.Generator(np.random.PCG64(seed))
    w = np.ascontiguousarray(E.random_fr_std(nvars, seed + 100))
    cls = rng.integers(0, 5, size=nvars)
    w[cls <= 1] = 0
    ones = cls == 2
    w[ones] = 0
    w[ones, 0] = 1
    small = cls == 3
    w[small] = 0
    w[small, 0] = rng.integers(0, 1 << 16, size=int(small.sum()), dtype=np.uint64)
    w[0] = (1, 0, 0, 0)
    return w


MASK_R = 0x0A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A
MASK_S = 0x1B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B3B


def cpu_prove(zk, wit, nthreads=None):
    """oracle/g16_oracle_cpu.cpp: the C++ restatement of the reference CPU prover (the checker and the CPU baseline)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    cores = nthreads or oc.ncpu()
    t0 = time.perf_counter()
    pa, pb, pc, phases = oc.prove(zk, wit, MASK_R, MASK_S, nthreads=cores)
    return (pa, pb, pc), phases, time.perf_counter() - t0, cores


def check_proof_against_ground_truth(g, zk, wit, raw, cpu_proof=None):
    """Ground truth for a proof of the benchmark fixture with the benchmark masks: (1) byte equality with the proof the
    compiled CPU restatement of the reference prover (prover.nim:215-304) computes from the same zkey / witness /
    masks, (2) the Groth16 verification equation of verifier.nim:31-52 (pairing restatement in the oracle)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bn254_pairing as bp
    from g16b200 import encoding as E
    if cpu_proof is None:
        cpu_proof = cpu_prove(zk, wit)[0]
    got = (np.frombuffer(bytes(raw.pi_a), dtype="<u8"), np.frombuffer(bytes(raw.pi_b), dtype="<u8"),
           np.frombuffer(bytes(raw.pi_c), dtype="<u8"))
    same = all(np.array_equal(a, np.asarray(b).reshape(-1)) for a, b in zip(got, cpu_proof))
    pub = E.fr_from_std(np.ascontiguousarray(wit[: zk.npubs + 1]))
    ok = bp.verify_proof(E.g1_from_array(zk.alpha1)[0], E.g2_from_array(zk.beta2)[0], E.g2_from_array(zk.gamma2)[0],
                         E.g2_from_array(zk.delta2)[0], E.g1_from_array(zk.pointsIC), pub,
                         E.g1_from_array(got[0])[0], E.g2_from_array(got[1])[0], E.g1_from_array(got[2])[0])
    if not same:
        log("PARITY FAILURE: the GPU proof differs from the CPU restatement of the reference prover")
    if not ok:
        log("PARITY FAILURE: the GPU proof does not satisfy the pairing equation")
    return bool(same and ok)


# ------------------------------------------------------------------------------------------------ reference arm
def fixture_via_child(log_n: int):
    """The reference arm must not run (or even map) this repo's CUDA library: the fixture -- a fake-setup zkey, for
    which the reference itself has no writer -- is generated by a child process on the GPU, written as snarkjs
    .zkey / .wtns files, and parsed back here with numpy views."""
    from g16b200 import files
    d = tempfile.mkdtemp(prefix="g16bench_")
    z, w = os.path.join(d, "c.zkey"), os.path.join(d, "c.wtns")
    subprocess.check_call([sys.executable, os.path.abspath(__file__), "--make-fixture", d, "--log-n", str(log_n)],
                          stdout=sys.stderr)
    zk = files.parse_zkey(z)
    wt = files.parse_witness(w)
    return zk, wt.values, d


def child_make_fixture(args):
    import g16b200 as g
    from g16b200 import files
    from g16b200.zkey_types import Witness
    zk, wit, _ = make_fixture(g, args.log_n)
    files.write_zkey(os.path.join(args.make_fixture, "c.zkey"), zk)
    files.write_witness(os.path.join(args.make_fixture, "c.wtns"), Witness(values=wit))


def run_reference(args):
    """The reference's CPU prover (prover.nim:215-304) as restated in oracle/g16_oracle_cpu.cpp -- the reference
    itself is Nim + un-vendored constantine and cannot be built here.  Each step is one full proof of
    the REAL 2^log_n fixture (the same zkey / witness / masks as the GPU arm) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sl = args.log_n if args.sample_log_n is None else min(args.sample_log_n, args.log_n)
    zk, wit, tmpdir = fixture_via_child(sl)
    times, cores = [], None
    for i in range(args.warmup + args.steps):
        _, _, dt, cores = cpu_prove(zk, wit)
        if i >= args.warmup:
            times.append(dt)
    import shutil
    shutil.rmtree(tmpdir, ignore_errors=True)
    per = sum(times) / len(times)
    scale = float(1 << (args.log_n - sl))
    value = 1.0 / (per * scale)
    sample = "one full CPU proof of the 2^%d fixture per step (%.2f s each)" % (sl, per)
    if scale != 1:
        sample += ", scaled x%d to 2^%d (extrapolated: --sample-log-n)" % (int(scale), args.log_n)
    out = {"impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * scale * 1e3,
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
            "parallelism": "1 gpu" if world == 1 else "MSM-level shard plan x%d (g16_shard_plan), one process per GPU, "
                                                      "NCCL all-gather of 400-byte records" % world,
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
    # a second communicator for the witness all-gather of the scatter upload mode (see ShardedProver.partials_submit_host)
    wgroup = dist.new_group() if dist is not None else None

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
    w_pinned_rows = w_pinned.view(-1, 4)
    w_dev = w_pinned.to("cuda")
    d2h_bytes = 256

    # `depth` proofs in flight (default 2): consecutive proofs overlap, so the latency-bound tail of one
    # (bucket-reduction levels, assembly, the all-gather) hides behind the accumulation kernels of the next.
    # Each proof in flight owns a context slot; the slots share one resident key (g16_ctx_clone).  depth 1 =
    # strictly sequential proofs, reported as `sequential`.
    depth = max(1, args.pipeline)
    trusted = not args.validate          # the fixture comes from our own setup: validation is timed in cold_e2e
    if world == 1:
        ctx = g.ProverContext(zk, trusted=trusted)
        ctxs = [ctx] + [ctx.clone() for _ in range(depth - 1)]

        def make_runner(ptr, mem_kind, d=None):
            def run(steps):
                dd = d or depth
                last = None
                for i in range(steps):
                    c = ctxs[i % dd]
                    if i >= dd:
                        last = c.wait()[0]
                    c.submit(ptr, mask, E.FORM_STD, mem_kind)
                for i in range(min(dd, steps)):
                    last = ctxs[(steps - min(dd, steps) + i) % dd].wait()[0]
                return last
            return run
    else:
        sp0 = g.parallel.ShardedProver(zk, rank, world, device=local, trusted=trusted)
        sps = [sp0] + [g.parallel.ShardedProver(zk, rank, world, device=local, share=sp0) for _ in range(depth - 1)]
        ctxs = [sp.ctx for sp in sps]
        ctx = ctxs[0]

        def make_runner(ptr, mem_kind, d=None):
            # per proof and rank: partial sums -> NCCL all-gather -> assembly, all enqueued without a host
            # synchronisation (ShardedProver.exchange_submit); the host only waits for the finished proof.
            # Host witness: every rank uploads 1/N of it and the slices are all-gathered over NVLink.
            def run(steps):
                dd = d or depth
                last = None
                for i in range(steps):
                    sp = sps[i % dd]
                    if i >= dd:
                        last = sp.wait()
                    if mem_kind == MEM_HOST and scatter_mode[0]:
                        sp.partials_submit_host(w_pinned_rows, mask, group=wgroup)
                    else:
                        sp.partials_submit(ptr, mem_kind, mask)
                    sp.exchange_submit(mask)
                for i in range(min(dd, steps)):
                    last = sps[(steps - min(dd, steps) + i) % dd].wait()
                return last
            return run

    scatter_mode = [world > 1 and not args.no_scatter]
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
    h2d_bytes = ctx.last_witness_bytes()               # what this rank actually copied for its last proof
    e2e_alt = None
    if scatter_mode[0]:
        # two ways to bring a host witness to N ranks: (a) each rank uploads 1/N and the slices are all-gathered over
        # NVLink (one more collective per proof), (b) each rank uploads the intervals its shard reads.  Both are timed,
        # the faster one is the e2e figure, the other is reported beside it.
        a_bytes = sps[0].h2d_bytes
        scatter_mode[0] = False
        b_dev_ms, b_wall_ms, _, raw2b = timed(run_e2e, args.steps, max(depth, args.warmup // 2))
        assert bytes(raw2b.pi_c) == bytes(raw2.pi_c)
        b_bytes = ctx.last_witness_bytes()
        if b_wall_ms < e2e_wall_ms:
            e2e_alt = {"mode": "scatter 1/N per rank + NVLink all-gather", "proofs_per_s": args.steps / (e2e_wall_ms * 1e-3)}
            e2e_dev_ms, e2e_wall_ms, h2d_bytes = b_dev_ms, b_wall_ms, b_bytes
            e2e_mode = "every rank uploads the witness intervals its shard reads"
        else:
            e2e_alt = {"mode": "every rank uploads the intervals its shard reads", "proofs_per_s": args.steps / (b_wall_ms * 1e-3)}
            h2d_bytes = a_bytes
            e2e_mode = "every rank uploads 1/N of the witness over its own PCIe link, NVLink all-gather of the slices"
    else:
        e2e_mode = "one upload of the witness" if world == 1 else "every rank uploads the witness intervals its shard reads"
    seq = None
    seq_dev_ms, seq_wall_ms, _, raw3 = timed(make_runner(w_dev.data_ptr(), MEM_DEVICE, 1), args.steps, 2)
    stats = dict(ctx.last_stats or {})                 # phases of a strictly sequential proof on this rank
    assert bytes(raw3.pi_c) == bytes(raw.pi_c)
    seq = {"ms_per_proof": seq_dev_ms / args.steps, "proofs_per_s": args.steps / (seq_dev_ms * 1e-3)}
    assert bytes(raw.pi_c) == bytes(raw2.pi_c) and bytes(raw.pi_a) == bytes(raw2.pi_a)

    # per-rank phases (device milliseconds of one sequential proof): maximum over ranks and the slowest rank
    phase = {k: round(v, 3) for k, v in stats.items() if k.startswith("ms_")}
    phase_info = {"max_over_ranks": phase, "slowest_rank": 0}
    total_h2d = h2d_bytes
    if dist is not None:
        allp = [None] * world
        mine = dict(phase)
        mine["h2d_bytes"] = h2d_bytes
        mine["plan"] = g.parallel.shard_plan(zk.nvars, zk.npubs, zk.domainSize, rank, world)
        dist.all_gather_object(allp, mine)
        keys = [k for k in phase]
        phase_info = {"max_over_ranks": {k: max(p[k] for p in allp) for k in keys},
                      "slowest_rank": max(range(world), key=lambda r: max(allp[r]["ms_msm_b2"], allp[r]["ms_msm_h"] +
                                                                           allp[r]["ms_quotient"] + allp[r]["ms_abc"],
                                                                           allp[r]["ms_msm_g1_witness"] +
                                                                           allp[r]["ms_sort_witness"])),
                      "per_rank": [{k: p[k] for k in keys + ["h2d_bytes"]} for p in allp],
                      "plan_fraction_of_each_array": [
                          {nm: [round(p["plan"][nm + "_lo"] / (zk.domainSize if nm == "h" else zk.nvars), 3),
                                round(p["plan"][nm + "_hi"] / (zk.domainSize if nm == "h" else zk.nvars), 3)]
                           for nm in ("h", "a1", "b1", "c1", "b2") if p["plan"][nm + "_hi"] > p["plan"][nm + "_lo"]}
                          for p in allp]}
        total_h2d = sum(p["h2d_bytes"] for p in allp)

    out = None
    parity_ok = True
    if rank == 0:
        value = args.steps / (dev_ms * 1e-3)
        out = {"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "u32x8-montgomery", "data": "synthetic",
               "config": dict(workload_config(args, world), proofs_in_flight=depth,
                              point_validation="off in the timed contexts (G16_ZKEY_TRUSTED): the fixture is this "
                                               "repo's own setup; the validated load is what cold_e2e times"
                              if trusted else "on"),
               "wall_ms_per_step": wall_ms / args.steps, "sequential": seq,
               "e2e": {"value": args.steps / (e2e_wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": total_h2d,
                       "d2h_bytes_per_step": d2h_bytes, "device_ms_per_step": e2e_dev_ms / args.steps,
                       "wall_ms_per_step": e2e_wall_ms / args.steps,
                       "note": "pinned host witness; " + e2e_mode + " (h2d_bytes_per_step = sum over ranks)",
                       "other_upload_mode": e2e_alt,
                       "api": "g16_prove_submit/wait (host witness)" if world == 1 else
                              "g16_ctx_set_mask + g16_prove_partials_submit + g16_ctx_order_stream + NCCL all-gather + "
                              "g16_prove_finish_submit + g16_prove_wait"},
               "gpu_launches": int(launches), "clocks": clocks,
               "phase_ms_last_step": phase_info}

    # ------------------------------------------------------------------ ground truth (outside the timed region)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            sl = min(args.cpu_sample_log_n, args.log_n)
            if sl == args.log_n:
                cpu_proof, phases, dt, cores = cpu_prove(zk, wit)
                parity_ok = check_proof_against_ground_truth(g, zk, wit, raw, cpu_proof)
                out["parity_checked"] = bool(parity_ok)
                out["parity"] = ("proof of the timed path == CPU restatement of the reference prover on the same "
                                 "fixture (byte equality) and passes the pairing verifier")
            else:
                zk_s, wit_s, _ = make_fixture(g, sl)
                cpu_proof, phases, dt, cores = cpu_prove(zk_s, wit_s)
                out["parity_checked"] = False
            scale = float(1 << (args.log_n - sl))
            labels = ["building 'ABC'", "computing the quotient (FFTs)", "computing pi_A (G1 MSM)",
                      "computing rho (G1 MSM)", "computing pi_B (G2 MSM)", "computing pi_C (2x G1 MSM)"]
            cpu = {"value": 1.0 / (dt * scale), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "one full CPU proof at 2^%d constraints (%.2f s)%s; C++ restatement of the reference "
                             "decomposition (constantine unavailable)" %
                             (sl, dt, "" if scale == 1 else ", scaled x%d to 2^%d" % (int(scale), args.log_n)),
                   "phase_seconds_sample": {k: round(v, 4) for k, v in zip(labels, phases)}}
        except Exception as ex:      # the checker being unavailable must not kill the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % (ex,)}
            out["parity_checked"] = False
        out["cpu_baseline"] = cpu

    # ------------------------------------------------------------------ extras (rank 0, N = 1)
    if world == 1 and not args.no_micro:
        out.update(micro_benchmarks(args, g, lib, zk, w_dev, torch))
        out.update(extra_lines(args, g, lib, zk, w_np, ctxs, mask, torch))
    for c in ctxs:
        c.close()
    if dist is not None:
        _lib.check(lib.g16_release_cached_memory())    # every rank gives its cached device memory back first
        torch.cuda.synchronize()
        dist.barrier()
        # rank 0 alone, every other rank idle: the SAME N GPUs driven by one process through the in-library
        # multi-device context (g16_ctx_create with shard_count = -N) -- the path a Nim caller of
        # generateProofWithMask gets; peer copies instead of NCCL
        # (the other ranks wait on the rendezvous store, not in an NCCL barrier whose kernel would spin on their GPUs)
        from datetime import timedelta
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            if not args.no_micro:
                try:
                    out["in_library_multi_gpu"] = inlib_line(args, g, lib, zk, w_pinned, w_dev, mask, raw, world, torch)
                except Exception as ex:
                    out["in_library_multi_gpu"] = {"failed": repr(ex)}
            store.set("g16_inlib_done", "1")
        else:
            store.wait(["g16_inlib_done"], timedelta(seconds=600))
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)
        if not parity_ok:
            sys.exit(3)


def inlib_line(args, g, lib, zk, w_pinned, w_dev, mask, raw_ref, world, torch):
    from g16b200 import _lib, encoding as E
    from g16b200.prover import MEM_DEVICE, MEM_HOST
    t0 = time.perf_counter()
    ctx = g.ProverContext(zk, devices=world, trusted=True)
    create_s = time.perf_counter() - t0
    slots = [ctx, ctx.clone()]
    res = {}
    for name, ptr, kind in (("resident", w_dev.data_ptr(), MEM_DEVICE), ("e2e", w_pinned.data_ptr(), MEM_HOST)):
        def run(steps):
            last = None
            for i in range(steps):
                c = slots[i % 2]
                if i >= 2:
                    last = c.wait()[0]
                c.submit(ptr, mask, E.FORM_STD, kind)
            for i in range(min(2, steps)):
                last = slots[(steps - min(2, steps) + i) % 2].wait()[0]
            return last
        run(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        raw = run(args.steps)
        dt = time.perf_counter() - t0
        assert bytes(raw.pi_a) == bytes(raw_ref.pi_a) and bytes(raw.pi_c) == bytes(raw_ref.pi_c)
        res[name] = {"proofs_per_s": args.steps / dt, "ms_per_proof": dt / args.steps * 1e3}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        slots[0].submit(w_dev.data_ptr(), mask, E.FORM_STD, MEM_DEVICE)
        slots[0].wait()
    res["sequential_ms_per_proof"] = (time.perf_counter() - t0) / args.steps * 1e3
    res["witness_bytes_per_proof"] = ctx.last_witness_bytes()
    res["e2e"]["witness"] = "pinned host memory, uploaded once in %d slices (one per device) + NVLink peer copies" % world
    # the other way to bring a host witness in: every device uploads the intervals its shard reads itself
    os.environ["G16_WITNESS_SCATTER"] = "0"
    try:
        alt = g.ProverContext(zk, devices=world, trusted=True)
        aslots = [alt, alt.clone()]

        def run2(steps):
            last = None
            for i in range(steps):
                c = aslots[i % 2]
                if i >= 2:
                    last = c.wait()[0]
                c.submit(w_pinned.data_ptr(), mask, E.FORM_STD, MEM_HOST)
            for i in range(min(2, steps)):
                last = aslots[(steps - min(2, steps) + i) % 2].wait()[0]
            return last
        run2(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        raw = run2(args.steps)
        dt = time.perf_counter() - t0
        assert bytes(raw.pi_c) == bytes(raw_ref.pi_c)
        res["e2e_direct_upload"] = {"proofs_per_s": args.steps / dt, "ms_per_proof": dt / args.steps * 1e3,
                                    "witness_bytes_per_proof": alt.last_witness_bytes()}
        for c in aslots:
            c.close()
    finally:
        del os.environ["G16_WITNESS_SCATTER"]
    res["context_create_s"] = create_s
    res["api"] = "g16_ctx_create(zkey, 0, -%d) + g16_prove_submit/wait: one process, %d devices, peer copies" % (world, world)
    res["same_proof_as_nccl_path"] = True
    for c in slots:
        c.close()
    return res


def extra_lines(args, g, lib, zk, w_np, ctxs, mask, torch):
    """Cold path, the drop-in call with a pageable Montgomery witness, and the skewed witness."""
    import numpy as np
    from g16b200 import _lib, encoding as E
    res = {}
    # (1) cold_e2e: what a caller of the reference's generateProofWithMask pays (cli_main.nim:193-210): context from
    # HOST zkey arrays (validated upload) + one proof + destroy, wall clock; one-shot layout vs resident tables
    for name, one_shot in (("cold_e2e", True), ("cold_e2e_tables", False)):
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = g.ProverContext(zk, one_shot=one_shot)
            t1 = time.perf_counter()
            raw, _ = c.prove_ptr(w_np.ctypes.data, mask, E.FORM_STD)
            t2 = time.perf_counter()
            c.close()
            t3 = time.perf_counter()
            ts.append((t3 - t0, t1 - t0, t2 - t1))
        best = min(ts)
        res[name] = {"ms": best[0] * 1e3, "create_ms": best[1] * 1e3, "prove_ms": best[2] * 1e3, "unit": "ms",
                     "first_call_ms": ts[0][0] * 1e3,
                     "layout": "plain points (G16_ZKEY_ONE_SHOT)" if one_shot else "resident window tables",
                     "point_validation": "on", "witness": "pageable host memory",
                     "what": "g16_ctx_create from host zkey arrays + g16_prove + g16_ctx_destroy, wall clock, best of 4 in one "
                             "process (device memory of a destroyed context stays in the library's pool); first_call_ms "
                             "is the first of them"}
    warm = 1e3 / res_value(ctxs, w_np, mask, E, 3)
    d_create = res["cold_e2e_tables"]["create_ms"] - res["cold_e2e"]["create_ms"]
    d_prove = res["cold_e2e"]["prove_ms"] - warm
    res["cold_e2e"]["crossover_proofs_per_key"] = (d_create / d_prove) if d_prove > 0 else None
    res["cold_e2e"]["crossover_note"] = ("tables cost %.0f ms more to build and save %.1f ms per proof: resident "
                                         "tables pay off from that many proofs per key" % (d_create, d_prove))
    # (2) the drop-in call as INTEGRATION.md binds it: synchronous g16_prove, pageable witness in Montgomery form
    c = ctxs[0]
    w_m = mont_witness(w_np)
    for _ in range(2):
        c.prove_ptr(w_m.ctypes.data, mask, E.FORM_MONT)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rawm, _ = c.prove_ptr(w_m.ctypes.data, mask, E.FORM_MONT)
    dt = (time.perf_counter() - t0) / args.steps
    res["e2e_dropin"] = {"value": 1.0 / dt, "unit": UNIT, "ms_per_proof": dt * 1e3,
                         "api": "g16_prove, synchronous, one proof at a time, witness = pageable Nim seq[Fr] payload "
                                "(Montgomery form) against a resident context"}
    # (3) skewed witness: timing of the same proof job on 40/20/20/20 scalars
    ws = skewed_witness(zk.nvars)
    wsp = torch.from_numpy(ws.view(np.int64).copy()).pin_memory()
    for _ in range(2):
        c.prove_ptr(wsp.data_ptr(), mask, E.FORM_STD)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.prove_ptr(wsp.data_ptr(), mask, E.FORM_STD)
    dts = (time.perf_counter() - t0) / args.steps
    wp = torch.from_numpy(w_np.view(np.int64).copy()).pin_memory()
    c.prove_ptr(wp.data_ptr(), mask, E.FORM_STD)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.prove_ptr(wp.data_ptr(), mask, E.FORM_STD)
    dtu = (time.perf_counter() - t0) / args.steps
    res["skewed"] = {"value": 1.0 / dts, "unit": UNIT, "ms_per_proof": dts * 1e3, "uniform_same_call_ms": dtu * 1e3,
                     "witness": "40 % zeros, 20 % ones, 20 % below 2^16, 20 % uniform (SURVEY.md 8d); timing only, "
                                "sequential g16_prove from pinned memory",
                     "phase_ms": {k: round(v, 3) for k, v in (c.last_stats or {}).items() if k.startswith("ms_")}}
    return res


def res_value(ctxs, w_np, mask, E, steps):
    t0 = time.perf_counter()
    for _ in range(steps):
        ctxs[0].prove_ptr(w_np.ctypes.data, mask, E.FORM_STD)
    return steps / (time.perf_counter() - t0)


def mont_witness(w_np):
    """The witness as a Nim seq[Fr] holds it: Montgomery residues (x * 2^256 mod r) in pageable host memory."""
    import numpy as np
    from g16b200 import encoding as E
    out = np.empty_like(w_np)
    step = 1 << 16
    Rm = (1 << 256) % E.R
    for i in range(0, w_np.shape[0], step):
        xs = E.fr_from_std(w_np[i:i + step])
        out[i:i + step] = E.ints_to_limbs([(x * Rm) % E.R for x in xs])
    return out


def micro_benchmarks(args, g, lib, zk, w_dev, torch):
    """Standalone G1 / G2 MSM 2^k and Fr NTT 2^k (BASELINE.json configs[2]) with inputs resident in HBM, plus the
    rooflines of the dominant kernels (MSM bucket accumulation in G1 and G2) and of the NTT passes."""
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

    result = torch.zeros(64, dtype=torch.int64, device="cuda")
    acc_ms, tot_ms, pairs = C.c_float(), C.c_float(), C.c_uint64()

    def run_msm(g2: bool, table: bool, pts):
        plan = C.c_void_p()
        _lib.check(lib.g16_msm_plan_create(1 if g2 else 0, n, 0, C.byref(plan)))
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

    pts = torch.from_numpy(zk.pointsA1.view("int64").copy()).to("cuda")
    res["msm_g1_plain"] = run_msm(False, False, pts)
    res["msm_g1"] = run_msm(False, True, pts)          # the layout the resident prover context uses
    del pts
    pts2 = torch.from_numpy(zk.pointsB2.view("int64").copy()).to("cuda")
    res["msm_g2"] = run_msm(True, True, pts2)
    del pts2
    # ncu --set full captures of the same kernels at 2^20 (profiles/, per-launch DRAM bytes read + written)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass

    def traffic(key):
        v = ncu.get(key)
        return (v.get("dram_bytes") if (v and log_n == v.get("log_n", 20)) else None), (v or {}).get("source")

    acc = res["msm_g1"]["bucket_accumulate_ms"]
    achieved = res["msm_g1"]["pairs"] * 10.0 * 136.0 / (acc * 1e-3)       # XYZZ mixed add = 8M + 2S
    tr, src = traffic("k_bucket_accumulate_g1")
    res["roofline"] = {"kernel": "k_bucket_accumulate<Fp> (G1 MSM over the resident window table, XYZZ mixed adds)",
                       "bound": "imad", "achieved": achieved / 1e12, "peak": mac_peak / 1e12, "unit": "TMAC32/s",
                       "frac": achieved / mac_peak,
                       "traffic": tr, "traffic_unit": "bytes per launch (ncu --set full, 2^20)", "traffic_source": src,
                       "algorithmic_bytes": res["msm_g1"]["pairs"] * 68.0,
                       "fmaheavy_cycles_active_pct_ncu": (ncu.get("k_bucket_accumulate_g1") or {}).get("fmaheavy_pct"),
                       "peak_source": "measured live: g16_bench_int_pipe(kind=3) x 136 MAC32 per Montgomery multiply "
                                      "(IMAD.WIDE.U32 is half rate; raw mad.lo.u32 rate %.2f T/s); independent "
                                      "hardware counter beside it: sm__pipe_fmaheavy_cycles_active from ncu"
                                      % (madlo_peak / 1e12),
                       "algorithmic_work": "pairs x 10 modmul x 136 MAC32"}
    acc2 = res["msm_g2"]["bucket_accumulate_ms"]
    ach2 = res["msm_g2"]["pairs"] * 28.0 * 136.0 / (acc2 * 1e-3)         # Fp2: 8 mul x 3 + 2 sqr x 2 Fp multiplies
    tr, src = traffic("k_bucket_accumulate_g2")
    res["roofline_g2"] = {"kernel": "k_bucket_accumulate<Fp2> (G2 MSM over the resident window table)",
                          "bound": "imad", "achieved": ach2 / 1e12, "peak": mac_peak / 1e12, "unit": "TMAC32/s",
                          "frac": ach2 / mac_peak, "traffic": tr, "traffic_source": src,
                          "algorithmic_bytes": res["msm_g2"]["pairs"] * 132.0,
                          "fmaheavy_cycles_active_pct_ncu": (ncu.get("k_bucket_accumulate_g2") or {}).get("fmaheavy_pct"),
                          "algorithmic_work": "pairs x 28 Fp modmul x 136 MAC32 (Karatsuba Fp2 multiply = 3, square = 2; "
                                              "a lazily reduced Fp2 multiply does fewer MAC32 than this accounting)"}
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
    passes = 1 if log_n <= 12 else 2 if log_n <= 24 else 3       # the algorithmic minimum
    alg_bytes = 64.0 * nn * passes
    ntt_modmul = nn / 2 * log_n
    tr, src = traffic("k_ntt_pass")
    res["ntt_fr"] = {"n": nn, "ms": t, "melem_per_s": nn / t / 1e3, "passes_algorithmic": passes}
    res["roofline_ntt"] = {"kernel": "k_ntt_pass (forward NTT 2^%d)" % log_n, "bound": "hbm",
                           "achieved": alg_bytes / (t * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": alg_bytes / (t * 1e-3) / 1e9 / hbm_peak,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                           "traffic": tr, "traffic_unit": "bytes per transform (ncu --set full, sum of its passes, 2^20)",
                           "traffic_source": src, "algorithmic_bytes": alg_bytes,
                           "imad_frac": ntt_modmul * 136.0 / (t * 1e-3) / mac_peak,
                           "note": "254-bit NTT is integer-pipe bound (SURVEY.md 8d); imad_frac is the binding roof"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 of the constraint count (headline: 20)")
    ap.add_argument("--sample-log-n", type=int, default=None,
                    help="--impl reference: prove a smaller instance per step and scale (default: the real size)")
    ap.add_argument("--cpu-sample-log-n", type=int, default=20,
                    help="cpu_baseline of the main arm: constraints (log2) of the one CPU proof that is timed")
    ap.add_argument("--pipeline", type=int, default=2, help="proofs in flight (1 = strictly sequential proofs)")
    ap.add_argument("--validate", action="store_true", help="validate the points in the timed contexts too")
    ap.add_argument("--no-scatter", action="store_true", help="N > 1: every rank copies the witness intervals it reads itself")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--make-fixture", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.make_fixture:
        child_make_fixture(args)
        return
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
