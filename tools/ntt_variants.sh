#!/bin/bash
# A/B of NTT tile configurations (libvar_<tilelog>_<threads>_<minblocks>.so built by hand from ntt.cu)
for lib in nim-groth16_b200/libg16b200.so nim-groth16_b200/libvar_10_128_7.so nim-groth16_b200/libvar_10_128_6.so nim-groth16_b200/libvar_10_256_3.so; do
  echo "== $lib"
  for lg in 16 20 22 24; do G16B200_LIB=$PWD/$lib timeout 300 python tools/ntt_probe.py $lg; done
  G16B200_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_core.py -x -q -m gpu -k "ntt" 2>&1 | tail -1
  G16B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-micro 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('bench', d['ms_per_step'], d['value'], d.get('sequential'))"
done
