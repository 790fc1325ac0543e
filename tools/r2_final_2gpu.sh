#!/bin/bash
# two GPUs: the multi-device and NCCL tests on real devices, bench at N = 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -x -q -m gpu > gpurun_out/r2_pytest_multi_2gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_multi_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2_l20.json 2> gpurun_out/r2h_bench_n2_l20.err; echo "n2 rc=$?"
tail -3 gpurun_out/r2h_bench_n2_l20.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2h_bench_n2_l20.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["note"][:80], d["e2e"].get("other_upload_mode"))
print(json.dumps(d.get("in_library_multi_gpu"))[:900])
PY
