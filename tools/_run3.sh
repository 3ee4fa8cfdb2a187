mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -s -k "fp16_accumulated or gv16_helpers" > gpurun_out/t_gv16.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_gv16.log
grep -v "^$" gpurun_out/t_gv16.log | grep -i "accumulated\|passed\|failed\|error\|exit" | cut -c1-400
for cfg in "tsa_rows fp32" "tsa_rows f16" "sca_rig fp32" "sca_rig mixed1" "sca_rig mixed2"; do
  set -- $cfg
  BENCH_GV=$2 BENCH_TAG=_gv$2 timeout 300 python tools/bench_msda.py --only $1 --iters 15 2>&1 | grep bfloat16 | cut -c1-250
done > gpurun_out/bench_gv.log 2>&1
cat gpurun_out/bench_gv.log
timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_gvoff.json 2> gpurun_out/bench_gvoff.err; tail -c 250 gpurun_out/bench_gvoff.json
BEVF_TSA_GV=f16 timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_tsaf16.json 2> gpurun_out/bench_tsaf16.err; tail -c 250 gpurun_out/bench_tsaf16.json
BEVF_TSA_GV=f16 BEVF_SCA_GV=mixed timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_bothf16.json 2> gpurun_out/bench_bothf16.err; tail -c 250 gpurun_out/bench_bothf16.json
BEVF_TSA_GV=f16 BEVF_SCA_GV=mixed timeout 900 python -m pytest tests/test_encoder_gpu.py -x -q -k "bf16" > gpurun_out/t_enc_gv16.log 2>&1; tail -3 gpurun_out/t_enc_gv16.log
