mkdir -p gpurun_out
nproc
timeout 600 python tools/flaky_hunt2.py 60 > gpurun_out/flaky_hunt2.log 2> gpurun_out/flaky_hunt2.err; cat gpurun_out/flaky_hunt2.log; tail -3 gpurun_out/flaky_hunt2.err
