#!/bin/bash
# multi-GPU check: N ranks via torchrun, bench at log_n (default 20)
N=${1:-2}; L=${2:-20}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --log-n $L > gpurun_out/bench_n${N}_l${L}.json 2> gpurun_out/bench_n${N}_l${L}.err
echo "rc=$?"; tail -3 gpurun_out/bench_n${N}_l${L}.err; cat gpurun_out/bench_n${N}_l${L}.json
