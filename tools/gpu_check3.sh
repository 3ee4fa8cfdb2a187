#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Timeout" | cut -c1-300 | tail -20
timeout 300 python tools/bench_msda.py --iters 10 --only rig 2>&1 | cut -c1-260
timeout 300 python tools/bench_gemm.py --iters 10 2>&1 | tee gpurun_out/bench_gemm.txt | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1e.json | cut -c1-1800
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 6 -o gpurun_out/prof_gemm_r1e python tools/bench_gemm.py --profile > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -c 2 -o gpurun_out/prof_msda_bwd_r1e python tools/bench_msda.py --profile --only sca_rig > gpurun_out/ncu_msda_bwd.log 2>&1; tail -2 gpurun_out/ncu_msda_bwd.log
