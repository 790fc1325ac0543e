#!/bin/bash
# round 2, two GPUs: the multi-device tests on real devices, the NCCL path against ground truth, bench at N = 2 for
# the line and the uniform shard policies, the single-GPU bench with the pooled allocator (cold path)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/smi2.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_pytest_multi2.log 2>&1; tail -15 gpurun_out/r2_pytest_multi2.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2_bench_n2_line.json 2> gpurun_out/r2_bench_n2_line.err; echo "n2 line rc=$?"
G16_SHARD_POLICY=uniform timeout 900 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 10 --warmup 4 --no-micro > gpurun_out/r2_bench_n2_uniform.json 2> gpurun_out/r2_bench_n2_uniform.err; echo "n2 uniform rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
for f in r2_bench_n2_line r2_bench_n2_uniform r2_bench2; do echo "== $f"; tail -2 gpurun_out/$f.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
    print(json.dumps(d.get("phase_ms_last_step"))[:1500])
    print(json.dumps(d.get("in_library_multi_gpu"))[:800])
    print(json.dumps(d.get("cold_e2e"))[:600])
except Exception as e: print("no json", e)
PY
done
