#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prover.py tests/test_gpu_multi.py tests/test_cpp_host.py -q -m gpu -k "not full_size" > gpurun_out/r2_pytest_glv.log 2>&1; tail -3 gpurun_out/r2_pytest_glv.log
for glv in 1 0; do
G16_GLV=$glv timeout 600 python bench.py --no-micro --no-cpu-baseline > gpurun_out/r2_bench_glv$glv.json 2>/dev/null
G16_GLV=$glv timeout 600 python bench.py --log-n 16 --no-micro --no-cpu-baseline > gpurun_out/r2_bench_glv${glv}_l16.json 2>/dev/null
python - <<PY
import json
for f in ("r2_bench_glv$glv","r2_bench_glv${glv}_l16"):
    d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
    print("GLV=$glv", f, {k:d.get(k) for k in ("value","ms_per_step")}, "seq", d["sequential"]["ms_per_proof"])
PY
echo "GLV=$glv shard probe"; G16_GLV=$glv PROBE_RANKS=2,4 timeout 600 python tools/shard_probe2.py 20 8 1 2 2>&1 | grep rank
done
