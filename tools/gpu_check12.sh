#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260 | sed 's/^/interleave=1 /'
BEVF_TSA_INTERLEAVE=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260 | sed 's/^/interleave=0 /'
done
for t in 8 16 32 64 128; do BEVF_CPU_THREADS=$t timeout 300 python - <<PY
import os, time, sys
sys.path.insert(0, '.')
import bench
step, q, w = bench.cpu_reference_step_factory(1, 1)
step(); t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
print('cpu threads', bench.cpu_threads(), 'sample s', round(dt, 2), 'q/s', round(q / dt, 1), flush=True)
PY
done
