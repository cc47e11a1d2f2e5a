mkdir -p gpurun_out
: > gpurun_out/repeat.jsonl
for lib in default v11 A; do
  if [ $lib = default ]; then unset XMIMSIM_B200_LIB; else export XMIMSIM_B200_LIB=$PWD/xmimsim_b200/lib/exp/lib_$lib.so; fi
  timeout 300 python tools/repeat_check.py 60 >> gpurun_out/repeat.jsonl 2>> gpurun_out/repeat.err
done
cat gpurun_out/repeat.jsonl; tail -3 gpurun_out/repeat.err
