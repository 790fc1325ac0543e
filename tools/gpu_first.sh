#!/bin/bash
# first GPU contact: smoke, int-pipe peaks, parity suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python - > gpurun_out/intpipe.log 2>&1 <<'PY'
import sys, ctypes as C
sys.path.insert(0,'nim-groth16_b200')
import g16b200
lib=g16b200._lib.load()
for kind,name in ((0,'mad.lo'),(1,'mad.hi'),(2,'wide pairs'),(3,'fmul')):
    ops=C.c_double(); ms=C.c_float()
    rc=lib.g16_bench_int_pipe(kind,C.byref(ops),C.byref(ms))
    print(name, rc, "%.4g ops/s"%ops.value, "%.3f ms"%ms.value, flush=True)
PY
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log; cat gpurun_out/intpipe.log; tail -40 gpurun_out/pytest_gpu.log
