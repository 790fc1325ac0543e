#!/bin/bash
# batched-affine tree mode: correctness (all GPU tests through the C ABI) and timing against the XYZZ item path
mkdir -p gpurun_out
for t in 5; do
  G16_MSM_TREE=$t timeout 300 python tools/msm_probe.py 20 0
  G16_MSM_TREE=$t timeout 300 python tools/msm_probe.py 20 1
done
G16_MSM_TREE=5 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for t in 0 4 5; do
  echo "== bench tree=$t"
  G16_MSM_TREE=$t timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d.get('sequential'))"
done
