mkdir -p gpurun_out
for lib in A B default; do
  if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
  (timeout 300 python -m pytest tests/test_history_gpu.py -m gpu -x -q -k "synthetic or batch_formation or digest") > gpurun_out/tests_bisect_$lib.log 2>&1; echo $lib rc=$?; tail -2 gpurun_out/tests_bisect_$lib.log
done
