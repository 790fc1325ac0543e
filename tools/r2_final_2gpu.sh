#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -x -q -m gpu > gpurun_out/r2_pytest_multi12.log 2>&1; tail -5 gpurun_out/r2_pytest_multi12.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29731 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2_bench_n2_l20_v2.json 2> gpurun_out/r2_bench_n2_l20_v2.err; echo "n2 rc=$?"
tail -3 gpurun_out/r2_bench_n2_l20_v2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n2_l20_v2.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
print(json.dumps(d.get("in_library_multi_gpu"))[:900])
PY
