mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:xmb_history -c 1 -f -o gpurun_out/hist_v12 python tools/bench_kernel.py 2000000 > gpurun_out/ncu_full_v12.log 2>&1; tail -3 gpurun_out/ncu_full_v12.log; ls -la gpurun_out/*.ncu-rep
