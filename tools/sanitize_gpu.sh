#!/bin/bash
# compute-sanitizer passes over the history kernel on small inputs (one gpurun call, a few minutes):
# initcheck (reads of uninitialised global memory: queues, accumulators), racecheck (shared-memory hazards: staging area,
# queue counters, batch sort, off-grid solid-angle rounds), memcheck.  Logs under gpurun_out/sanitize_*.log.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
n=${1:-4000}
for tool in initcheck racecheck memcheck; do
  for w in syn srm; do
    timeout ${2:-50} $CS --tool $tool --print-limit 20 python tools/sanitize_gpu_run.py $w $n > gpurun_out/sanitize_${tool}_$w.log 2>&1
    echo "$tool $w rc=$? $(grep -c 'hazard\|Uninitialized\|Invalid' gpurun_out/sanitize_${tool}_$w.log) reports; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_$w.log | tail -1)"
  done
done
