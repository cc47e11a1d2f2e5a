mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_history_gpu.py tests/test_brute_gpu.py tests/test_advanced_compton_gpu.py tests/test_escape_gpu.py -m gpu -x -q) > gpurun_out/tests_v14d.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/tests_v14d.log
timeout 200 python tools/bench_brute.py 4000000 --no-cpu > gpurun_out/brute_v14.json 2> gpurun_out/brute_v14.err; cat gpurun_out/brute_v14.json
: > gpurun_out/ab_v14d.jsonl
timeout 200 python tools/bench_kernel.py 2000000 srm1412 >> gpurun_out/ab_v14d.jsonl 2>> gpurun_out/ab_v14d.err
timeout 200 python tools/bench_kernel.py 20000000 synthetic10 >> gpurun_out/ab_v14d.jsonl 2>> gpurun_out/ab_v14d.err
cat gpurun_out/ab_v14d.jsonl
