#!/bin/bash
# 8 GPUs, after the all-pairs peer access fix: bench N = 8 at 2^20 (in-library e2e) and the reference arm under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29771 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n8_l20.json 2> gpurun_out/r2g_bench_n8_l20.err; echo "n8 l20 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29772 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/r2g_bench_ref_n8.json 2> gpurun_out/r2g_bench_ref_n8.err; echo "ref n8 rc=$?"; grep "^{" gpurun_out/r2g_bench_ref_n8.json | cut -c1-200
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2g_bench_n8_l20.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"].get("other_upload_mode"))
print(json.dumps(d.get("in_library_multi_gpu"))[:1500])
PY
