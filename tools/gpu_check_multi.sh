#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | tee gpurun_out/bench_r1j_n$N.json | cut -c1-900
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --no-graph 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | cut -c1-400
