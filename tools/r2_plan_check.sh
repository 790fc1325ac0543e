#!/bin/bash
# per-rank busy times of the planned shards (two proofs in flight) on one GPU
mkdir -p gpurun_out
{
for G in 2 3 4; do timeout 300 python tools/shard_probe2.py 20 $G 2 2>&1 | grep "^rank"; done
PROBE_RANKS=0,2,5 timeout 300 python tools/shard_probe2.py 22 8 2 2>&1 | grep "^rank"
timeout 300 python tools/shard_probe2.py 22 4 2 2>&1 | grep "^rank"
timeout 300 python tools/shard_probe2.py 16 8 2 2>&1 | grep "^rank"
timeout 300 python tools/shard_probe2.py 16 2 2 2>&1 | grep "^rank"
} | tee gpurun_out/r2_plan_check2.log
