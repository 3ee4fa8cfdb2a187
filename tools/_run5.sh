mkdir -p gpurun_out
# ncu: full capture of the mixed-accumulation backward kernel (SCA launch) + the fp16 TSA kernel
BENCH_GV=mixed2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_bwd_d32 -c 1 -o gpurun_out/ncu_bwd_mixed python tools/bench_msda.py --only sca_rig --profile > gpurun_out/ncu_bwd_mixed.log 2>&1
ncu -i gpurun_out/ncu_bwd_mixed.ncu-rep --page raw --csv > gpurun_out/ncu_bwd_mixed_raw.csv 2>/dev/null
# launch list of two bench steps (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-graph --no-standin --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-200
# other configs + A/B
timeout 600 python bench.py --config small --no-standin --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 200 gpurun_out/bench_small.json
timeout 600 python bench.py --config tiny --no-standin --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; tail -c 200 gpurun_out/bench_tiny.json
BEVF_GEMM=cublas timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_cublas.json 2> gpurun_out/bench_cublas.err; tail -c 300 gpurun_out/bench_cublas.json; tail -3 gpurun_out/bench_cublas.err
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; tail -12 gpurun_out/bench_gemm.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
