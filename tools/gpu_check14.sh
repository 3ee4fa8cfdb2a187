#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -c 2 -o gpurun_out/prof_msda_bwd_r1p python tools/bench_msda.py --profile --only sca_rig > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/launches_r1p.csv python bench.py --no-graph --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log | cut -c1-200
