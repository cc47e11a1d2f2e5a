mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_history_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q) > gpurun_out/tests_v14c.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/tests_v14c.log
timeout 300 python tools/bench_configs.py > gpurun_out/configs_v14.json 2>gpurun_out/configs_v14.err; cat gpurun_out/configs_v14.json
XMB_LAYER_SORT=0 timeout 300 python tools/bench_configs.py > gpurun_out/configs_v14_sort0.json 2>>gpurun_out/configs_v14.err; cat gpurun_out/configs_v14_sort0.json
: > gpurun_out/ab_v14c.jsonl
for lib in default u1 u4 b1 b7 default; do
  if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
  timeout 200 python tools/bench_kernel.py 2000000 srm1412 >> gpurun_out/ab_v14c.jsonl 2>> gpurun_out/ab_v14c.err
done
cat gpurun_out/ab_v14c.jsonl
