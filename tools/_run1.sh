mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_dense_gpu.py -x -q -s > gpurun_out/t_dense.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_dense.log
tail -12 gpurun_out/t_dense.log
for cfg in "0 3" "1 3" "2 3" "10 3" "4 3" "0 1" "1 1" "2 1" "10 1"; do
  set -- $cfg
  echo "== debug=$1 teams=$2"
  BENCH_DENSE=1 BEVF_DENSE_TEAMS=$2 BEVF_DENSE_DEBUG=$1 BENCH_TAG=_dbg timeout 300 python tools/bench_msda.py --only sca_rig --iters 8 --kernels 2>&1 | grep "bfloat16" | grep dense_tc
done > gpurun_out/bench_dense_dbg.log 2>&1
cat gpurun_out/bench_dense_dbg.log
BENCH_DENSE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_bwd_dense -c 1 -o gpurun_out/ncu_dense python tools/bench_msda.py --only sca_rig --profile > gpurun_out/ncu_dense.log 2>&1
tail -3 gpurun_out/ncu_dense.log
ncu -i gpurun_out/ncu_dense.ncu-rep --page raw --csv > gpurun_out/ncu_dense_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_dense.ncu-rep --page source --csv > gpurun_out/ncu_dense_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
