mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xmb_history -c 1 -f -o gpurun_out/hist_v13 python tools/bench_kernel.py 2000000 > gpurun_out/ncu_full_v13.log 2>&1; tail -3 gpurun_out/ncu_full_v13.log
