#!/bin/bash
# eight GPUs: bench at N = 8 for 2^20 (line and uniform policies) and 2^22 (line and "G2 on its own GPU"), the
# multi-GPU MSM / NTT sweep
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_l20.json 2> gpurun_out/r2_bench_n8_l20.err; echo "n8 l20 rc=$?"
G16_SHARD_POLICY=uniform timeout 900 $TR --nproc-per-node 8 --master-port 29722 bench.py --gpus 8 --steps 20 --warmup 5 --no-micro --no-cpu-baseline > gpurun_out/r2_bench_n8_l20_uniform.json 2> gpurun_out/r2_bench_n8_l20_uniform.err; echo "n8 l20 uniform rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29723 bench.py --gpus 8 --steps 8 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2_bench_n8_l22.json 2> gpurun_out/r2_bench_n8_l22.err; echo "n8 l22 rc=$?"
G16_SHARD_POLICY=g2own timeout 900 $TR --nproc-per-node 8 --master-port 29724 bench.py --gpus 8 --steps 8 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2_bench_n8_l22_g2own.json 2> gpurun_out/r2_bench_n8_l22_g2own.err; echo "n8 l22 g2own rc=$?"
PARITY_MAX=16 timeout 900 $TR --nproc-per-node 8 --master-port 29725 tools/sweep_multi.py 12 24 > gpurun_out/r2_sweep_n8.jsonl 2> gpurun_out/r2_sweep_n8.err; echo "sweep rc=$?"; tail -3 gpurun_out/r2_sweep_n8.jsonl | cut -c1-400
for f in r2_bench_n8_l20 r2_bench_n8_l20_uniform r2_bench_n8_l22 r2_bench_n8_l22_g2own; do echo "== $f"; tail -1 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
    ph=d.get("phase_ms_last_step") or {}
    for r,p in enumerate(ph.get("per_rank",[])): print("  rank",r,{k:v for k,v in p.items() if k not in ("ms_h2d","ms_assemble")}, ph["plan_fraction_of_each_array"][r])
    print(json.dumps(d.get("in_library_multi_gpu"))[:900])
except Exception as e: print("no json", e)
PY
done
