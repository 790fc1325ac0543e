#!/bin/bash
# ncu launch list of the fixed workload on the final code (after the run without ncu has exited 0)
mkdir -p gpurun_out
timeout 300 python tools/profile_run.py 20 > gpurun_out/r2h_profile_run_plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_profile_run_2p20.csv python tools/profile_run.py 20 > gpurun_out/r2h_profile_run_ncu.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/r2h_profile_run_plain.log
