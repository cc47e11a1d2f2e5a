#!/bin/bash
# Parity hunt with host-side contention: W concurrent workers, each R fresh processes of the history parity tests up to
# and including the 10-layer test.  Mismatches are described in gpurun_out/parity_mismatch_<pid>.json (tests/helpers.py).
# usage: tools/parity_hunt2.sh W R
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-4}; R=${2:-25}
worker() {
  local w=$1
  for i in $(seq 1 $R); do
    timeout 600 python -m pytest tests/test_history_gpu.py -x -q -m gpu -p no:cacheprovider -k "examples or caso4 or option or synthetic_ten" 2>&1 | tail -1 | sed "s/^/w$w r$i: /" >> gpurun_out/parity_hunt2.log
  done
}
: > gpurun_out/parity_hunt2.log
for w in $(seq 1 $W); do worker $w & done
wait
echo "passed runs: $(grep -c ' passed' gpurun_out/parity_hunt2.log)  failed runs: $(grep -c 'failed' gpurun_out/parity_hunt2.log)"
ls gpurun_out/parity_mismatch_* 2>/dev/null
