#!/bin/bash
# A/B of experiment builds (tools/build_variant.sh) on the two kernel shapes: usage tools/ab.sh NAME...   ("default" = the built library)
cd "$(dirname "$0")/.."
for n in "$@"; do
  if [ "$n" = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$n.so; fi
  python tools/bench_kernel.py 2000000 | sed "s/^/$n: /"
  python tools/bench_kernel.py 20000000 synthetic10 | sed "s/^/$n: /"
done
