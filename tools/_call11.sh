mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_history_gpu.py -m gpu -x -q) > gpurun_out/tests_v14a.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/tests_v14a.log
: > gpurun_out/ab_v14.jsonl
for m in 0 3 0 3; do
  XMB_LAYER_SORT=$m timeout 200 python tools/bench_kernel.py 2000000 srm1412 >> gpurun_out/ab_v14.jsonl 2>> gpurun_out/ab_v14.err
done
for m in 0 3; do
  XMB_LAYER_SORT=$m timeout 200 python tools/bench_kernel.py 150000 srm1155 >> gpurun_out/ab_v14.jsonl 2>> gpurun_out/ab_v14.err
done
cat gpurun_out/ab_v14.jsonl
