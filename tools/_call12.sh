mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_history_gpu.py -m gpu -x -q) > gpurun_out/tests_v14b.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/tests_v14b.log
: > gpurun_out/ab_v14b.jsonl
for m in 15 30 7 15 30; do
  XMB_ECLS_MAX=$m timeout 200 python tools/bench_kernel.py 2000000 srm1412 >> gpurun_out/ab_v14b.jsonl 2>> gpurun_out/ab_v14b.err
done
for lib in epl0 default epl0 default; do
  if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
  timeout 200 python tools/bench_kernel.py 20000000 synthetic10 >> gpurun_out/ab_v14b.jsonl 2>> gpurun_out/ab_v14b.err
done
cat gpurun_out/ab_v14b.jsonl
