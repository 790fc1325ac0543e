#!/bin/bash
mkdir -p gpurun_out
for r in 5 2 0; do
PROBE_RANKS=$r timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_rank$r.csv python tools/shard_probe2.py 20 8 1 > gpurun_out/r2_launches_rank$r.log 2>&1; tail -1 gpurun_out/r2_launches_rank$r.log
done
