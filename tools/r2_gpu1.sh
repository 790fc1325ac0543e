#!/bin/bash
# round 2, first GPU pass: smoke, the new multi-device / validation tests, the whole GPU suite, the default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_pytest_multi.log 2>&1; tail -15 gpurun_out/r2_pytest_multi.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/r2_pytest_all.log 2>&1; tail -5 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench1.err
cut -c1-3000 gpurun_out/r2_bench1.json
