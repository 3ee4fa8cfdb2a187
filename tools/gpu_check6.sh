#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_r1h_graph.json | cut -c1-2200
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph 2>&1 | tail -1 | tee gpurun_out/bench_r1h_eager.json | cut -c1-400
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Timeout" | cut -c1-300 | tail -8
