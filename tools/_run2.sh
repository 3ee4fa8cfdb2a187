mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_gpu_full.log
tail -5 gpurun_out/t_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --no-standin > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
timeout 600 python bench.py --no-standin --no-cpu-baseline --dense 1 > gpurun_out/bench_dense1.json 2> gpurun_out/bench_dense1.err; tail -c 300 gpurun_out/bench_dense1.json
BEVF_GEMM=cublas timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_cublas.json 2> gpurun_out/bench_cublas.err; tail -c 300 gpurun_out/bench_cublas.json
