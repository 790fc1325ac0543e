#!/bin/bash
G16_MSM_TREE=5 timeout 900 python -m pytest tests/test_gpu_core.py -x -q -m gpu -k "msm" 2>&1 | tail -3
for t in 5 4; do
  G16_MSM_TREE=$t timeout 300 python tools/msm_probe.py 20 0
  G16_MSM_TREE=$t timeout 300 python tools/msm_probe.py 20 1
done
G16_MSM_TREE=5 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tree_launches_g1.csv python tools/msm_probe.py 20 0 > gpurun_out/tree_ncu.log 2>&1
G16_MSM_TREE=5 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tree_launches_g2.csv python tools/msm_probe.py 20 1 > gpurun_out/tree_ncu2.log 2>&1
