#!/bin/bash
# round-2 ncu evidence on the final code + the single-GPU sweep
mkdir -p gpurun_out
timeout 600 python tools/profile_run.py 20 > gpurun_out/r2f_profile_run_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_profile_run_2p20.csv python tools/profile_run.py 20 > gpurun_out/r2f_profile_run_ncu.log 2>&1
echo "launch list rc=$?"
timeout 600 python tools/msm_probe.py 20 1 > gpurun_out/r2f_msm_probe_g2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bucket_accumulate -c 2 -o gpurun_out/r2f_g2_acc python tools/msm_probe.py 20 1 > gpurun_out/r2f_ncu_g2.log 2>&1
echo "ncu g2 rc=$?"
timeout 600 python tools/msm_probe.py 20 0 > gpurun_out/r2f_msm_probe_g1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bucket_accumulate -c 2 -o gpurun_out/r2f_g1_acc python tools/msm_probe.py 20 0 > gpurun_out/r2f_ncu_g1.log 2>&1
echo "ncu g1 rc=$?"
SWEEP_CPU_MAX=18 timeout 1200 python tools/sweep.py 12 24 > gpurun_out/r2f_sweep_1gpu.jsonl 2> gpurun_out/r2f_sweep_1gpu.err; echo "sweep rc=$?"
tail -2 gpurun_out/r2f_sweep_1gpu.jsonl | cut -c1-300
