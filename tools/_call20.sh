mkdir -p gpurun_out
timeout 500 python tools/flaky_hunt.py 80 > gpurun_out/flaky_hunt.json 2> gpurun_out/flaky_hunt.err; cat gpurun_out/flaky_hunt.json; tail -3 gpurun_out/flaky_hunt.err
