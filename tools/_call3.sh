mkdir -p gpurun_out
timeout 200 python tools/limbs_digest.py > gpurun_out/digest_new2.json 2>&1
XMB_LAYER_SORT=1 timeout 200 python tools/limbs_digest.py > gpurun_out/digest_new2_ls1.json 2>&1
python - <<'PY'
import json
a=json.load(open('tools/_digest_old.json'))
for nm in ('new2','new2_ls1'):
    try: b=json.load(open('gpurun_out/digest_%s.json'%nm))
    except Exception as e: print(nm, 'FAILED', open('gpurun_out/digest_%s.json'%nm).read()[-2000:]); continue
    for k in a: print(nm, k, 'SAME' if (a[k]['sha'],a[k]['sha_shard'])==(b[k]['sha'],b[k]['sha_shard']) else 'DIFF', a[k]['ms'], b[k]['ms'])
PY
: > gpurun_out/variants3.jsonl
for ls in 2 1; do
  XMB_LAYER_SORT=$ls timeout 120 python tools/bench_kernel.py 20000000 synthetic10 >> gpurun_out/variants3.jsonl 2>gpurun_out/var3_err.log
done
timeout 120 python tools/bench_kernel.py 2000000 srm1412 >> gpurun_out/variants3.jsonl 2>>gpurun_out/var3_err.log
cat gpurun_out/variants3.jsonl | cut -c1-220; tail -3 gpurun_out/var3_err.log
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_v12.log 2>&1; echo tests rc=$?; tail -2 gpurun_out/tests_v12.log
