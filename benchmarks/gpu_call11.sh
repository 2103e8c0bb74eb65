#!/bin/bash
mkdir -p gpurun_out
export MR_BENCH_SKIP_CPU=1
for kb in 1 8 16 32 64; do
  MR_WGRAD_MIN_KB=$kb timeout 120 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/s3i_c3_kb$kb.json 2> gpurun_out/s3i_c3_kb$kb.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s3i_c3_kb$kb.json').read().strip().splitlines()[-1])
    print('kb=$kb', d['ms_per_step'], d['value'], d['final_loss'])
except Exception as e:
    print('kb=$kb', 'failed', e)
PY
done
