#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cpu_bound_check.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k wgrad 2>&1 | grep -E "ERR|assert|passed|failed|Error" | cut -c1-300 | tail -6
timeout 300 python tools/bench_gemm.py --iters 10 2>&1 | tee gpurun_out/bench_gemm_ws2.txt | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1g.json | cut -c1-500
