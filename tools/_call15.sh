mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_escape_gpu.py -m gpu -x -q) > gpurun_out/tests_v14e.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_v14e.log
timeout 300 python tools/bench_escape.py --no-cpu > gpurun_out/escape_v14.json 2> gpurun_out/escape_v14.err; cat gpurun_out/escape_v14.json; tail -3 gpurun_out/escape_v14.err
