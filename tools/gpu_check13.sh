#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 4 -o gpurun_out/prof_msda_final python tools/bench_msda.py --profile --only sca_rig > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 6 -o gpurun_out/prof_gemm_final python tools/bench_gemm.py --profile > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r1n_default.json | cut -c1-2600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_r1n_reference.json | cut -c1-700
