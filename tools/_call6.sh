mkdir -p gpurun_out
: > gpurun_out/ab_v12b.jsonl
for w in "2000000 srm1412" "20000000 synthetic10"; do
  for lib in default b7 b4 b1 b0 b0u4 default; do
    if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
    timeout 200 python tools/bench_kernel.py $w >> gpurun_out/ab_v12b.jsonl 2>> gpurun_out/ab_v12b.err
  done
done
cat gpurun_out/ab_v12b.jsonl
export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_b0.so
(timeout 700 python -m pytest tests/test_history_gpu.py -m gpu -x -q -k "digest or bit or batch_formation or examples") > gpurun_out/tests_b0.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/tests_b0.log
