#!/bin/bash
for p in lmh hmh hlm hhh mlh hml; do
  echo "== prio $p"
  G16_STREAM_PRIO=$p timeout 300 python tools/shard_probe.py 20 8 2>&1 | tail -1 | cut -c1-120
  G16_STREAM_PRIO=$p timeout 300 python tools/shard_probe.py 20 1 2>&1 | tail -1 | cut -c1-120
done
