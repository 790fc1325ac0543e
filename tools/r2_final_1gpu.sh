#!/bin/bash
# round-2 single-GPU evidence on the final code
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2f_smi.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2f_pytest_gpu.log
G16_GRAPH=1 timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_multi.py -q -m gpu -k "not nccl and not full_size" > gpurun_out/r2f_pytest_graph.log 2>&1; tail -3 gpurun_out/r2f_pytest_graph.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "bench rc=$?"
G16_GRAPH=1 timeout 600 python bench.py --no-micro --no-cpu-baseline > gpurun_out/r2f_bench_1gpu_graph.json 2> /dev/null; echo "bench graph rc=$?"
timeout 600 python bench.py --log-n 16 --no-micro --cpu-sample-log-n 16 > gpurun_out/r2f_bench_1gpu_l16.json 2> /dev/null; echo "bench l16 rc=$?"
G16_GRAPH=1 timeout 600 python bench.py --log-n 16 --no-micro --no-cpu-baseline > gpurun_out/r2f_bench_1gpu_l16_graph.json 2> /dev/null; echo "bench l16 graph rc=$?"
timeout 600 python bench.py --log-n 22 --steps 4 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/r2f_bench_1gpu_l22.json 2> /dev/null; echo "bench l22 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("r2f_bench_1gpu","r2f_bench_1gpu_graph","r2f_bench_1gpu_l16","r2f_bench_1gpu_l16_graph","r2f_bench_1gpu_l22","r2f_bench_ref"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked","gpu_launches")}, (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "no json", e)
PY
