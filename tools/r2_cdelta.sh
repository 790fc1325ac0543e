#!/bin/bash
# window width A/B on one GPU: G16_C_DELTA shifts the window of every MSM of the context
mkdir -p gpurun_out
for d in 0 -1 -2; do
G16_C_DELTA=$d timeout 300 python bench.py --no-micro --no-cpu-baseline > gpurun_out/r2_cdelta_$d.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_cdelta_$d.json") if l.startswith("{")][-1])
print("C_DELTA=$d", {k:d.get(k) for k in ("value","ms_per_step")}, "seq", d["sequential"]["ms_per_proof"], "e2e", d["e2e"]["value"])
PY
done 2>&1 | tee gpurun_out/r2_cdelta.log
