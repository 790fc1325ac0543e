#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_prover.py tests/test_gpu_multi.py -x -q -m gpu -k "msm or shard or multi_device or golden or closed_form" > gpurun_out/r2_pytest_items.log 2>&1; tail -3 gpurun_out/r2_pytest_items.log
echo "== shard probe (new items)"; PROBE_RANKS=0,2,5,7 timeout 600 python tools/shard_probe2.py 20 8 1 2 2>&1 | grep rank | tee gpurun_out/r2_shard_probe_items.log
echo "== shard probe (old items)"; G16_MSM_TARGET_ITEMS=0 PROBE_RANKS=5 timeout 600 python tools/shard_probe2.py 20 8 1 2 2>&1 | grep rank
SWEEP_CPU_MAX=16 timeout 900 python tools/sweep.py 14 21 > gpurun_out/r2_sweep_1gpu_items.jsonl 2> gpurun_out/r2_sweep_1gpu_items.err
G16_MSM_TARGET_ITEMS=0 SWEEP_CPU_MAX=0 timeout 900 python tools/sweep.py 14 21 > gpurun_out/r2_sweep_1gpu_olditems.jsonl 2> /dev/null
python - <<PY
import json
for f in ("r2_sweep_1gpu_items","r2_sweep_1gpu_olditems"):
    print(f)
    for l in open("gpurun_out/%s.jsonl"%f):
        if not l.startswith("{"): continue
        d=json.loads(l)
        print("  ", d["log_n"], "g1", d["g1_table"]["ms"], d["g1_table"]["c"], d["g1_table"]["accumulate_frac_of_modmul_peak"], d.get("g1_cpu_matches"), "g2", d["g2_table"]["ms"], d["g2_table"]["accumulate_frac_of_modmul_peak"], d.get("g2_cpu_matches"), "agree", d["g1_layouts_agree"], d["g2_layouts_agree"])
PY
timeout 600 python bench.py --no-cpu-baseline --no-micro > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench11.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","sequential")}, d["e2e"]["value"])
PY
