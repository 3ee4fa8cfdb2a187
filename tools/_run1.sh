mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_dense_gpu.py -x -q -s > gpurun_out/t_dense.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_dense.log
tail -25 gpurun_out/t_dense.log
BEVF_DENSE_TILES=8 BEVF_DENSE_TEAMS=4 timeout 600 python -m pytest tests/test_msda_dense_gpu.py -x -q -s -k "equals_plain or stale" > gpurun_out/t_dense_8x4.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_dense_8x4.log
tail -8 gpurun_out/t_dense_8x4.log
for cfg in "1 16 3 512 8192 _t3" "2 16 3 512 8192 _t3s2" "1 16 2 512 8192 _t2" "2 16 2 512 8192 _t2s2" "1 8 4 512 8192 _t8x4" "1 16 3 1024 8192 _t3c1024" "1 16 3 512 2048 _t3coarse23" "2 16 2 512 2048 _t2s2coarse23" "1 16 1 512 8192 _t1"; do
  set -- $cfg
  echo "== dense=$1 tiles=$2 teams=$3 chunk=$4 maxpix=$5"
  BENCH_DENSE=$1 BEVF_DENSE_TILES=$2 BEVF_DENSE_TEAMS=$3 BEVF_DENSE_CHUNK=$4 BEVF_DENSE_MAXPIX=$5 BENCH_TAG=$6 timeout 300 python tools/bench_msda.py --only sca_rig --iters 15 --kernels 2>&1 | grep "bfloat16" | tail -3
done > gpurun_out/bench_dense.log 2>&1
cat gpurun_out/bench_dense.log
