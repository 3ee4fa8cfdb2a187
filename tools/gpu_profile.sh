#!/bin/bash
# The ncu recipes behind profiles/ (run through gpurun; always bounded with -k / -c / -s):
#   1. `--set full` of the sampler kernels on the real SCA geometry and of the tcgen05 GEMMs
#   2. launch list (gpu__time_duration) of ~1.2 eager training steps of bench.py
# Summaries are exported here with `ncu -i <rep> --page raw --csv` and committed under profiles/.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 6 \
    -o gpurun_out/prof_msda python tools/bench_msda.py --profile --only sca_rig > gpurun_out/ncu_msda.log 2>&1
tail -1 gpurun_out/ncu_msda.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 6 \
    -o gpurun_out/prof_gemm python tools/bench_gemm.py --profile > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv \
    --log-file gpurun_out/launches.csv python bench.py --no-graph --no-cpu-baseline --steps 2 --warmup 3 \
    > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-200
