mkdir -p gpurun_out
BENCH_GV=mixed2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_bwd_d32 --launch-skip 1 --launch-count 1 -o gpurun_out/ncu_bwd_mixed python tools/bench_msda.py --only sca_rig --profile > gpurun_out/ncu_bwd_mixed.log 2>&1
ncu -i gpurun_out/ncu_bwd_mixed.ncu-rep --page raw --csv > gpurun_out/ncu_bwd_mixed_raw.csv 2>/dev/null
head -c 400 gpurun_out/ncu_bwd_mixed_raw.csv | tail -c 200; tail -2 gpurun_out/ncu_bwd_mixed.log
