mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/tests_v11b.log 2>&1; echo tests rc=$?; tail -14 gpurun_out/tests_v11b.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; echo bench rc=$?; cat gpurun_out/bench_v11.json
timeout 300 python tools/bench_configs.py > gpurun_out/configs_v11.json 2>gpurun_out/configs_v11.err; cat gpurun_out/configs_v11.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v11.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xmb_history -c 1 -f -o gpurun_out/hist_v11 python tools/bench_kernel.py 2000000 > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:xmb_history -c 1 -f -o gpurun_out/hist_v11_syn python tools/bench_kernel.py 20000000 synthetic10 > gpurun_out/ncu_full_syn.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_v11.json 2>&1; cat gpurun_out/bench_ref_v11.json
