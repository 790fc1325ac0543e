#!/bin/bash
# 4 GPUs after the planner change: NCCL ranks against the ground truth, multi-device context, bench at N = 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "nccl or multi_device_context or recombine" > gpurun_out/r2g_pytest_multi_4gpu.log 2>&1; tail -3 gpurun_out/r2g_pytest_multi_4gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29761 bench.py --gpus 4 --steps 20 --warmup 5 --no-micro --no-cpu-baseline > gpurun_out/r2g_bench_n4_l20.json 2> gpurun_out/r2g_bench_n4_l20.err; echo "n4 l20 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2g_bench_n4_l20.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"].get("other_upload_mode"))
print(json.dumps(d.get("in_library_multi_gpu"))[:700])
print(json.dumps(d.get("phase_ms_last_step"))[:900])
PY
