#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1750 -c 640 --csv --log-file gpurun_out/launches_step_r1i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
echo "ncu rows: $(wc -l < gpurun_out/launches_step_r1i.csv)"
