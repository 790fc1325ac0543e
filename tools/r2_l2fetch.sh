#!/bin/bash
# L2 fetch granularity A/B (G16_L2_FETCH): accumulate time, DRAM bytes of the G1 accumulate kernel, whole proof
mkdir -p gpurun_out
{
for v in default 64 32; do
  if [ $v = default ]; then unset G16_L2_FETCH; else export G16_L2_FETCH=$v; fi
  echo "== L2_FETCH=$v"
  timeout 120 python tools/msm_probe.py 20 0 2>&1 | tail -1
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_bucket_accumulate --launch-skip 2 -c 1 --csv python tools/msm_probe.py 20 0 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  if [ $v != 32 ]; then
  timeout 200 python bench.py --no-micro --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('bench', d['value'], d['ms_per_step'], d['clocks'])"
  fi
done
} 2>&1 | tee gpurun_out/r2_l2fetch.log
