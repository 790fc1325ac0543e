#!/bin/bash
# full GPU validation + short bench + shard probe
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('sequential'), d.get('micro'))"
timeout 600 python tools/shard_probe.py 20 0 2>&1 | tail -4
