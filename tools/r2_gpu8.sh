#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_tail.log 2>&1; tail -4 gpurun_out/r2_pytest_tail.log
echo "== shard probe"; timeout 600 python tools/shard_probe2.py 20 8 1 2 3 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_tail.log
timeout 600 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench8.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential")}, d["e2e"]["value"])
PY
