#!/bin/bash
mkdir -p gpurun_out
for l1 in 16 8 4; do for l2 in 8192 2048; do echo "== L1=$l1 L2MIN=$l2"; G16_REDUCE_L1=$l1 G16_REDUCE_L2MIN=$l2 PROBE_RANKS=0,2,5 timeout 600 python tools/shard_probe2.py 20 8 1 2 2>&1 | grep rank; done; done | tee gpurun_out/r2_shard_probe_reduce.log
