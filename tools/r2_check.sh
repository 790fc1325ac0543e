#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 4 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2z_bench.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked","gpu_launches")}, d["e2e"]["value"], d["cold_e2e"]["ms"])
PY
