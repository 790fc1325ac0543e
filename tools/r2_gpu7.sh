#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 600 python tools/shard_probe2.py 20 8 1 2 3 4 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_default.log
for d in -1 -2 -3; do echo "== G16_C_DELTA=$d (ranks 2,5)"; G16_C_DELTA=$d PROBE_RANKS=0,2,5 timeout 600 python tools/shard_probe2.py 20 8 1 2 3 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_c$d.log; done
