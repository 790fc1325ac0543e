#!/bin/bash
# final 4-GPU evidence: N = 4 at 2^20 and 2^22, N = 2 at 2^22
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29751 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n4_l20.json 2> gpurun_out/r2f_bench_n4_l20.err; echo "n4 l20 rc=$?"
timeout 900 $TR --nproc-per-node 4 --master-port 29752 bench.py --gpus 4 --steps 6 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2f_bench_n4_l22.json 2> gpurun_out/r2f_bench_n4_l22.err; echo "n4 l22 rc=$?"
timeout 900 $TR --nproc-per-node 2 --master-port 29753 bench.py --gpus 2 --steps 6 --warmup 3 --log-n 22 --no-micro --no-cpu-baseline > gpurun_out/r2f_bench_n2_l22.json 2> gpurun_out/r2f_bench_n2_l22.err; echo "n2 l22 rc=$?"
for f in r2f_bench_n4_l20 r2f_bench_n4_l22 r2f_bench_n2_l22; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"].get("other_upload_mode"))
    print(json.dumps(d.get("in_library_multi_gpu"))[:1200])
except Exception as e: print("no json", e)
PY
done
