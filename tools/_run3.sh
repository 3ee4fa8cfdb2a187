mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -s -k "bf16_accumulated" > gpurun_out/t_gv16.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_gv16.log
grep -v "^$" gpurun_out/t_gv16.log | tail -8
for gv in fp32 bf16; do
  BENCH_GV=$gv BENCH_TAG=_gv$gv timeout 300 python tools/bench_msda.py --only tsa_rows --iters 15 2>&1 | grep bfloat16 | cut -c1-260
done > gpurun_out/bench_gv.log 2>&1
cat gpurun_out/bench_gv.log
BEVF_TSA_GV=bf16 timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_tsagv16.json 2> gpurun_out/bench_tsagv16.err; tail -c 400 gpurun_out/bench_tsagv16.json
BEVF_TSA_GV=bf16 timeout 900 python -m pytest tests/test_encoder_gpu.py -x -q -k "bf16" > gpurun_out/t_enc_gv16.log 2>&1; tail -3 gpurun_out/t_enc_gv16.log
