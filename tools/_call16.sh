mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_history_gpu.py tests/test_brute_gpu.py tests/test_escape_gpu.py tests/test_advanced_compton_gpu.py -m gpu -x -q) > gpurun_out/tests_v15a.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_v15a.log
timeout 300 python tools/bench_escape.py --no-cpu > gpurun_out/escape_v15.json 2> gpurun_out/escape_v15.err; cut -c1-300 gpurun_out/escape_v15.json
timeout 200 python tools/bench_brute.py 4000000 --no-cpu > gpurun_out/brute_v15.json 2> gpurun_out/brute_v15.err; cut -c1-300 gpurun_out/brute_v15.json
: > gpurun_out/ab_v15.jsonl
timeout 200 python tools/bench_kernel.py 2000000 srm1412 >> gpurun_out/ab_v15.jsonl 2>> gpurun_out/ab_v15.err
timeout 200 python tools/bench_kernel.py 20000000 synthetic10 >> gpurun_out/ab_v15.jsonl 2>> gpurun_out/ab_v15.err
cat gpurun_out/ab_v15.jsonl
