#!/bin/bash
# One GPU-box session: GEMM unit tests first (each case in a subprocess with a timeout), then the
# whole GPU suite, the headline bench with the tcgen05 GEMMs and with the library GEMMs (A/B), and a
# bounded ncu launch list of one steady-state step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k "test_linear_tc" > gpurun_out/gemm_nt.log 2>&1; NT=$?
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k "wgrad" > gpurun_out/gemm_wgrad.log 2>&1; WG=$?
grep -E "ERR|Error|error|assert|passed|failed" gpurun_out/gemm_nt.log | cut -c1-300 | tail -25
grep -E "ERR|Error|error|assert|passed|failed" gpurun_out/gemm_wgrad.log | cut -c1-300 | tail -25
if [ $NT -ne 0 ]; then export BEVF_GEMM=cublas; echo "NT GEMM FAILED -> library GEMMs for the rest"; fi
if [ $WG -ne 0 ]; then export BEVF_WGRAD=cublas; echo "WGRAD FAILED -> library wgrad for the rest"; fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gemm_gpu.py 2>&1 | grep -E "^E  |FAILED|passed|failed|Timeout" | cut -c1-330 | tail -40
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tc.json | cut -c1-1500
BEVF_GEMM=cublas timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cublas.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "ncu rows: $(wc -l < gpurun_out/launches_step.csv)"
