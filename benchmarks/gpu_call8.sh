#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_conv_engine_gpu.py -q -k "transpose" > gpurun_out/s3g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3g_tests.log
timeout 150 python bench.py --config 5 > gpurun_out/s3g_bench_c5.json 2> gpurun_out/s3g_bench_c5.err
MR_NO_ENGINE_CONVT=1 timeout 150 python bench.py --config 5 > gpurun_out/s3g_bench_c5_libconvt.json 2> gpurun_out/s3g_bench_c5_libconvt.err
tail -3 gpurun_out/s3g_tests.log
for f in s3g_bench_c5 s3g_bench_c5_libconvt; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'])
except Exception as e:
    print('$f', 'failed', e)
PY
done
