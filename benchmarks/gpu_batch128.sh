#!/bin/bash
mkdir -p gpurun_out
export MR_BENCH_SKIP_CPU=1
timeout 100 python bench.py --config 3 --batch 128 > gpurun_out/s3k_c3_b128.json 2> gpurun_out/s3k_c3_b128.err
timeout 100 python bench.py --config 4 --batch 128 > gpurun_out/s3k_c4_b128.json 2> gpurun_out/s3k_c4_b128.err
for f in s3k_c3_b128 s3k_c4_b128; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'])
except Exception as e:
    print('$f', 'failed', e)
PY
done
