#!/bin/bash
# ncu captures of the product's kernels (one gpurun call): --set full with source for the history kernel on the two
# shapes (srm1412: two-layer kernel; configs3: generic kernel with per-layer queues), the solid-angle kernel and the
# detector-response kernels.  Reports land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.
# usage: tools/profile_kernels.sh TAG [what...]   what: hist syn sa det (default: all)
cd "$(dirname "$0")/.."
tag=${1:-vX}; shift
what=${@:-hist syn sa det}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
for w in $what; do
  case $w in
    hist) timeout 400 $NCU -k regex:xmb_history -s 1 -c 1 -o gpurun_out/hist_$tag python tools/kernel_counters.py srm1412 2000000 2 > gpurun_out/ncu_hist_$tag.log 2>&1;;
    syn)  timeout 400 $NCU -k regex:xmb_history -s 1 -c 1 -o gpurun_out/hist_${tag}_syn python tools/kernel_counters.py configs3 20000000 2 > gpurun_out/ncu_syn_$tag.log 2>&1;;
    sa)   timeout 300 $NCU -k regex:xmb_solid_angle -c 1 -o gpurun_out/sa_$tag python tools/kernel_counters.py srm1132 1000 1 > gpurun_out/ncu_sa_$tag.log 2>&1;;
    det)  timeout 300 ncu --set full --clock-control none -f -k regex:'det_|escape_|pileup|convolve|prefix|poisson|response|counts_to' -c 40 -o gpurun_out/det_$tag python tools/detector_run.py > gpurun_out/ncu_det_$tag.log 2>&1;;
  esac
  echo "$w rc=$?"
done
ls -la gpurun_out/*.ncu-rep
