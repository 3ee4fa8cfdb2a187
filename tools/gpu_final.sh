#!/bin/bash
# round-end check: full GPU test suite, smoke(), default bench (with cpu_baseline), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r1r_default.json | cut -c1-1200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_r1r_reference.json | cut -c1-400
