#!/bin/bash
mkdir -p gpurun_out
echo "== G=4"; timeout 600 python tools/shard_probe2.py 20 4 2 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_g4.log
echo "== G=2"; timeout 600 python tools/shard_probe2.py 20 2 2 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_g2.log
echo "== G=8"; timeout 600 python tools/shard_probe2.py 20 8 2 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_g8.log
