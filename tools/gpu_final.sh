#!/bin/bash
# end-of-round evidence: smoke, GPU suite, bench (both arms), ncu launch list of the profile workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
timeout 600 python tools/profile_run.py 20 > gpurun_out/profile_run_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tools/profile_run.py 20 > gpurun_out/profile_run_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python bench.py --log-n 22 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_2p22.json 2> gpurun_out/bench_2p22.err; echo "2p22 rc=$?"
cat gpurun_out/bench_final.json | cut -c1-1500
