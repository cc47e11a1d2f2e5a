#!/bin/bash
# Parity hunt on the GPU box: fresh processes of (a) the driver's command up to the 10-layer test, (b) the history file
# alone, (c) the 10-layer test alone.  Every mismatch writes gpurun_out/parity_mismatch_<pid>.json (tests/helpers.py).
# usage: tools/parity_hunt.sh N_DRIVER N_FILE N_SINGLE
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/parity_hunt.log
: > $LOG
nd=${1:-3}; nf=${2:-6}; ns=${3:-20}
for i in $(seq 1 $nd); do
  echo "== driver-like run $i" >> $LOG
  timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 >> $LOG
done
for i in $(seq 1 $nf); do
  echo "== history file run $i" >> $LOG
  timeout 600 python -m pytest tests/test_history_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $LOG
done
for i in $(seq 1 $ns); do
  echo "== 10-layer alone $i" >> $LOG
  timeout 300 python -m pytest tests/test_history_gpu.py -x -q -m gpu -k synthetic_ten 2>&1 | tail -2 >> $LOG
done
grep -c passed $LOG; grep -c failed $LOG
ls gpurun_out/parity_mismatch_* 2>/dev/null
