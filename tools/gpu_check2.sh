#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py tests/test_encoder_gpu.py -m gpu -q --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Timeout" | cut -c1-300 | tail -20
echo "== default (QR on, tile on)"; timeout 300 python tools/bench_msda.py --iters 10 --only rig 2>&1 | cut -c1-260
echo "== QR off"; BEVF_MSDA_QR=0 timeout 300 python tools/bench_msda.py --iters 10 --only sca_rig 2>&1 | cut -c1-260
echo "== iters 1"; BEVF_MSDA_ITERS=1 timeout 300 python tools/bench_msda.py --iters 10 --only sca_rig 2>&1 | cut -c1-260
echo "== tile off"; BENCH_TILE=0 timeout 300 python tools/bench_msda.py --iters 10 --only sca_rig 2>&1 | cut -c1-260
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1d.json | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 4 -o gpurun_out/prof_msda_r1d python tools/bench_msda.py --profile --only sca_rig > gpurun_out/ncu_r1d.log 2>&1; tail -2 gpurun_out/ncu_r1d.log
