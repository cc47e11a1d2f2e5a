#!/bin/bash
# compute-sanitizer over the PARITY TESTS themselves (real kernel-computed 128 x 128 grid, 30 000 photons of the
# 10-layer / 8-interaction sample incl. its off-grid solid-angle rounds, engine + oracle in the pytest process):
# initcheck (uninitialised global reads: queues, accumulators, grid), memcheck, racecheck (shared-memory hazards),
# synccheck (divergent barriers).  usage: tools/sanitize_gpu_tests.sh [per-tool timeout s] [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
T=${1:-240}
K=${2:-synthetic_ten}
for tool in initcheck memcheck synccheck racecheck; do
  log=gpurun_out/sanitize_tests_${tool}.log
  SECONDS=0; timeout $T $CS --tool $tool --print-limit 30 --target-processes all \
      python -m pytest tests/test_history_gpu.py -x -q -m gpu -p no:cacheprovider -k "$K" > $log 2>&1
  echo "$tool rc=$? ${SECONDS}s | $(grep -E 'passed|failed' $log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(tail -1 $log)"
done
