#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_engine_gpu.py tests/test_trunks_engine_gpu.py tests/test_surfaces_gpu.py -q > gpurun_out/s3e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3e_tests.log
for c in 3 4 5; do
  timeout 150 python bench.py --config $c > gpurun_out/s3e_bench_c$c.json 2> gpurun_out/s3e_bench_c$c.err
done
tail -15 gpurun_out/s3e_tests.log
for c in 3 4 5; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s3e_bench_c$c.json').read().strip().splitlines()[-1])
    print($c, d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['stages']['launch'])
except Exception as e:
    print($c, 'failed', e)
PY
done
