#!/bin/bash
# lazily reduced Fp2 multiplication: self-test, G2 parity tests, A/B timing of the G2 accumulate kernel, whole proof
mkdir -p gpurun_out
python - <<PY
import sys; sys.path.insert(0, "nim-groth16_b200")
from g16b200 import _lib
lib = _lib.load()
print("selftest rc", lib.g16_selftest(7, 4096), lib.g16_last_error())
PY
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_prover.py -x -q -m gpu -k "msm or golden or closed_form or pairing or selftest" > gpurun_out/r2_pytest_lazy.log 2>&1; tail -4 gpurun_out/r2_pytest_lazy.log
for lib in libg16b200.so libg16b200_nolazy.so; do
  for mb in 3 2; do
    echo "== $lib MINB=$mb"; G16B200_LIB=$PWD/nim-groth16_b200/$lib G16_G2_MINB=$mb timeout 300 python tools/msm_probe.py 20 1 2>&1 | tail -1
  done
done | tee gpurun_out/r2_g2_lazy_ab.log
timeout 600 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
G16B200_LIB=$PWD/nim-groth16_b200/libg16b200_nolazy.so timeout 600 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r2_bench4_nolazy.json 2> gpurun_out/r2_bench4_nolazy.err; echo "bench nolazy rc=$?"
python - <<PY
import json
for f in ("r2_bench4","r2_bench4_nolazy"):
    d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
    print(f, {k:d.get(k) for k in ("value","ms_per_step","sequential")}, d["e2e"]["value"])
PY
