#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k wgrad 2>&1 | grep -E "ERR|assert|passed|failed|Error" | cut -c1-300 | tail -6
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Timeout" | cut -c1-300 | tail -6
timeout 300 python tools/bench_gemm.py --iters 10 2>&1 | cut -c1-250
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1l_graph.json | cut -c1-330
