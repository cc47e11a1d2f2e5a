# One gpurun call at the end of a round: GPU test suite, bench line (+ reference arm), launch list, ncu full sets of the history
# kernel on both shapes.  usage: gpurun --timeout 2400 -- 'bash tools/gpu_round_end.sh TAG'   (then tools/ncu_summary.py / ncu_regions.py here)
tag=${1:-vX}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/tests_$tag.log 2>&1; echo tests rc=$?; tail -14 gpurun_out/tests_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench rc=$?; cut -c1-400 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ncu > gpurun_out/bench_under_ncu.log 2>&1
tools/profile_kernels.sh $tag hist syn
