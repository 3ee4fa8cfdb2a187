mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -k "fp16_accumulated or gv16_helpers" > gpurun_out/t_gv16.log 2>&1; tail -2 gpurun_out/t_gv16.log
BEVF_TSA_GV=f16 BEVF_SCA_GV=mixed timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_plan_gpu.py tests/test_transformer_gpu.py -x -q > gpurun_out/t_enc_gv16.log 2>&1; tail -3 gpurun_out/t_enc_gv16.log
BEVF_TSA_GV=f16 timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_tsaf16.json 2> gpurun_out/bench_tsaf16.err; tail -c 250 gpurun_out/bench_tsaf16.json
BEVF_TSA_GV=f16 BEVF_SCA_GV=mixed timeout 600 python bench.py --no-standin --no-cpu-baseline --breakdown gpurun_out/breakdown_bothf16.txt > gpurun_out/bench_bothf16.json 2> gpurun_out/bench_bothf16.err; tail -c 250 gpurun_out/bench_bothf16.json
BEVF_TSA_GV=f16 BEVF_SCA_GV=mixed BEVF_GV_MAXCONTRIB=64 timeout 600 python bench.py --no-standin --no-cpu-baseline > gpurun_out/bench_bothf16_c64.json 2> gpurun_out/bench_bothf16_c64.err; tail -c 250 gpurun_out/bench_bothf16_c64.json
head -30 gpurun_out/breakdown_bothf16.txt | cut -c1-150
