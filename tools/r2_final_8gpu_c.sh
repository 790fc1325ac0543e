#!/bin/bash
# 8-GPU box, final code: bench N = 8 and N = 4 at 2^20 (scatter upload mode on its own NCCL group)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29781 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n8_l20.json 2> gpurun_out/r2i_bench_n8_l20.err; echo "n8 rc=$?"
timeout 900 $TR --nproc-per-node 4 --master-port 29782 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n4_l20.json 2> gpurun_out/r2i_bench_n4_l20.err; echo "n4 rc=$?"
for f in r2i_bench_n8_l20 r2i_bench_n4_l20; do python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
print("$f", {k:d.get(k) for k in ("value","ms_per_step","sequential","parity_checked")}, d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["note"][:70], d["e2e"].get("other_upload_mode"))
print(json.dumps(d.get("in_library_multi_gpu"))[:700])
PY
done
