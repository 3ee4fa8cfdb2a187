#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Timeout" | cut -c1-300 | tail -10
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1m_graph.json | cut -c1-330
