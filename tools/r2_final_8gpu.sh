#!/bin/bash
# final 8-GPU evidence: bench N = 8 at 2^20 and 2^22, the multi-GPU sweep
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29741 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n8_l20.json 2> gpurun_out/r2f_bench_n8_l20.err; echo "n8 l20 rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29743 bench.py --gpus 8 --steps 8 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2f_bench_n8_l22.json 2> gpurun_out/r2f_bench_n8_l22.err; echo "n8 l22 rc=$?"
PARITY_MAX=16 timeout 900 $TR --nproc-per-node 8 --master-port 29745 tools/sweep_multi.py 12 24 > gpurun_out/r2f_sweep_n8.jsonl 2> gpurun_out/r2f_sweep_n8.err; echo "sweep rc=$?"
for f in r2f_bench_n8_l20 r2f_bench_n8_l22; do echo "== $f"; tail -1 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"].get("other_upload_mode"))
    ph=d.get("phase_ms_last_step") or {}
    for r,p in enumerate(ph.get("per_rank",[])): print("  rank",r,{k:v for k,v in p.items() if k not in ("ms_h2d","ms_assemble")}, ph["plan_fraction_of_each_array"][r])
    print(json.dumps(d.get("in_library_multi_gpu"))[:900])
except Exception as e: print("no json", e)
PY
done
