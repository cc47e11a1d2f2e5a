mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_history_gpu.py -m gpu -x -q) > gpurun_out/tests_v13b.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/tests_v13b.log
: > gpurun_out/ab_v13b.jsonl
for w in "2000000 srm1412" "20000000 synthetic10"; do
  for lib in v12 default s0p1 s1p0 c0 default; do
    if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
    timeout 200 python tools/bench_kernel.py $w >> gpurun_out/ab_v13b.jsonl 2>> gpurun_out/ab_v13b.err
  done
done
cat gpurun_out/ab_v13b.jsonl
