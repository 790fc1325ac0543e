#!/bin/bash
# 4 GPUs, 2^22 constraints, after the planner change
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 $TR --nproc-per-node 4 --master-port 29771 bench.py --gpus 4 --steps 8 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2g_bench_n4_l22.json 2> gpurun_out/r2g_bench_n4_l22.err; echo "n4 l22 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2g_bench_n4_l22.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"])
print(json.dumps(d.get("phase_ms_last_step"))[:600])
PY
