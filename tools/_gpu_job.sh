mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
tail -c 3000 gpurun_out/bench_v7.json
python tools/bench_escape.py > gpurun_out/escape_bench.json 2> gpurun_out/escape_bench.err
cat gpurun_out/escape_bench.json; tail -3 gpurun_out/escape_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:xmb_history_kernel -c 2 --csv --log-file gpurun_out/traffic_v7.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
tail -5 gpurun_out/traffic_v7.csv
