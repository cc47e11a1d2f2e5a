# One gpurun call: GPU test suite, bench line, the other BASELINE configs, launch list, ncu full sets, reference arm.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round_end.sh TAG'
tag=${1:-vX}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/tests_$tag.log 2>&1; echo tests rc=$?; tail -14 gpurun_out/tests_$tag.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench rc=$?; cat gpurun_out/bench_$tag.json
timeout 300 python tools/bench_configs.py > gpurun_out/configs_$tag.json 2>gpurun_out/configs_$tag.err; cat gpurun_out/configs_$tag.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xmb_history -c 1 -f -o gpurun_out/hist_$tag python tools/bench_kernel.py 2000000 > gpurun_out/ncu_full_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:xmb_history -c 1 -f -o gpurun_out/hist_${tag}_syn python tools/bench_kernel.py 20000000 synthetic10 > gpurun_out/ncu_full_syn_$tag.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2>&1; cat gpurun_out/bench_ref_$tag.json
cat gpurun_out/parity_retries.log 2>/dev/null
