#!/bin/bash
# The whole history parity file in N consecutive fresh processes ("30 consecutive fresh-process runs ... zero retries").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=${1:-30}; ok=0; bad=0
: > gpurun_out/history_runs.log
for i in $(seq 1 $n); do
  if timeout 600 python -m pytest tests/test_history_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/history_run_$i.log 2>&1; then ok=$((ok+1)); rm -f gpurun_out/history_run_$i.log; else bad=$((bad+1)); fi
  tail -1 gpurun_out/history_run_$i.log 2>/dev/null >> gpurun_out/history_runs.log
done
echo "test_history_gpu.py in fresh processes: $ok passed, $bad failed of $n" | tee -a gpurun_out/history_runs.log
