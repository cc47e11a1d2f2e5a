mkdir -p gpurun_out
for lib in default v11; do
  if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
  for i in 1 2 3 4 5 6 7 8; do
    (timeout 120 python -m pytest tests/test_history_gpu.py -m gpu -q -k "examples or caso4 or option or synthetic or continuous or gaussian" 2>&1 | tail -1) | sed "s/^/$lib $i: /"
  done
done 2>&1 | tee gpurun_out/flaky.log
