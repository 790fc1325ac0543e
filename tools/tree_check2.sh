#!/bin/bash
for t in 0 5 3; do
 for p in 2 4; do
  echo "== bench tree=$t pipeline=$p"
  G16_MSM_TREE=$t timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-micro --pipeline $p 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d.get('sequential'))"
 done
done
