#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/s3b_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/s3b_gputests.log
for c in 3 4 5; do
  timeout 150 python bench.py --config $c > gpurun_out/s3b_bench_c$c.json 2> gpurun_out/s3b_bench_c$c.err
done
tail -5 gpurun_out/s3b_gputests.log
for c in 3 4 5; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s3b_bench_c$c.json').read().strip().splitlines()[-1])
    print($c, d['ms_per_step'], d['value'], d['e2e']['value'], d['stages']['launch'])
except Exception as e:
    print($c, 'failed', e)
PY
done
