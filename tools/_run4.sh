mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_gpu_full.log
tail -4 gpurun_out/t_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --breakdown gpurun_out/breakdown_default.txt > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 700 gpurun_out/bench_default.json
BEVF_GV_ACC=fp32 timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_gvfp32.json 2> gpurun_out/bench_gvfp32.err; tail -c 250 gpurun_out/bench_gvfp32.json
