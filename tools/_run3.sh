mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -s -k "fp16_accumulated and sca" > gpurun_out/t_gv16_sca.log 2>&1
grep -i "accumulated\|passed\|failed" gpurun_out/t_gv16_sca.log | cut -c1-400
for cfg in "sca_rig fp32" "sca_rig f16" "sca_rig mixed1" "sca_rig mixed2" "sca_rig mixed3"; do
  set -- $cfg
  BENCH_GV=$2 BENCH_TAG=_gv$2 timeout 300 python tools/bench_msda.py --only $1 --iters 15 --kernels 2>&1 | grep bfloat16 | cut -c1-200
done > gpurun_out/bench_gv_sca.log 2>&1
cat gpurun_out/bench_gv_sca.log
BEVF_TSA_GV=f16 BEVF_SCA_GV=mixed BEVF_GV_MAXCONTRIB=100000 timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_allf16.json 2> gpurun_out/bench_allf16.err; tail -c 250 gpurun_out/bench_allf16.json
