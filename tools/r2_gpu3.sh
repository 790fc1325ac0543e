#!/bin/bash
# NTT radix-8 kernel: parity tests, A/B timing against the radix-2 kernel, ncu traffic, whole-proof effect
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_prover.py -x -q -m gpu -k "ntt or quotient or golden or closed_form or intermediates" > gpurun_out/r2_pytest_ntt.log 2>&1; tail -5 gpurun_out/r2_pytest_ntt.log
echo "== radix-8"; timeout 300 python tools/ntt_probe2.py 11 12 16 18 20 22 24 2>&1 | tee gpurun_out/r2_ntt_probe_radix8.log
echo "== radix-2"; G16_NTT_RADIX2=1 timeout 300 python tools/ntt_probe2.py 16 20 22 24 2>&1 | tee gpurun_out/r2_ntt_probe_radix2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass8 -c 4 -o gpurun_out/r2_ntt_pass8 python tools/ntt_probe2.py 20 > gpurun_out/r2_ncu_ntt.log 2>&1; tail -2 gpurun_out/r2_ncu_ntt.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench3.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential")}, d["e2e"]["value"], d.get("ntt_fr"), d.get("cold_e2e"))
PY
