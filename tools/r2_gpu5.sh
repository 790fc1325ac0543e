#!/bin/bash
# four GPUs: NCCL world-4 parity test, C++ host --gpus, bench at N = 4 (2^20 with the in-library line, 2^22), N = 2 at
# 2^22, and the reference arm on the real fixture (one step)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -x -q -m gpu -k "nccl or multi_gpu or in_library" > gpurun_out/r2_pytest_multi4.log 2>&1; tail -6 gpurun_out/r2_pytest_multi4.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29711 bench.py --gpus 4 --steps 10 --warmup 4 > gpurun_out/r2_bench_n4_l20.json 2> gpurun_out/r2_bench_n4_l20.err; echo "n4 l20 rc=$?"
timeout 900 $TR --nproc-per-node 4 --master-port 29712 bench.py --gpus 4 --steps 6 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2_bench_n4_l22.json 2> gpurun_out/r2_bench_n4_l22.err; echo "n4 l22 rc=$?"
timeout 900 $TR --nproc-per-node 2 --master-port 29713 bench.py --gpus 2 --steps 6 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2_bench_n2_l22.json 2> gpurun_out/r2_bench_n2_l22.err; echo "n2 l22 rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref1.json 2> gpurun_out/r2_bench_ref1.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2_bench_ref1.json
for f in r2_bench_n4_l20 r2_bench_n4_l22 r2_bench_n2_l22; do echo "== $f"; tail -2 gpurun_out/$f.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
    ph=d.get("phase_ms_last_step") or {}
    for r,p in enumerate(ph.get("per_rank",[])): print("  rank",r,p, ph["plan_fraction_of_each_array"][r])
    print(json.dumps(d.get("in_library_multi_gpu"))[:900])
except Exception as e: print("no json", e)
PY
done
