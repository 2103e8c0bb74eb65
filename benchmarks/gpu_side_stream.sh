#!/bin/bash
# weight gradients on a side stream (conv_engine.WGRAD_SIDE_STREAM): parity tests + A/B of the cfg-3 / cfg-4 steps
mkdir -p gpurun_out
export MR_BENCH_SKIP_CPU=1
timeout 200 python -m pytest tests/test_conv_engine_gpu.py tests/test_trunks_engine_gpu.py -q > gpurun_out/s3l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3l_tests.log
timeout 100 python bench.py --config 3 > gpurun_out/s3l_c3_side.json 2> gpurun_out/s3l_c3_side.err
MR_CONV_WGRAD_SIDE_STREAM=0 timeout 100 python bench.py --config 3 > gpurun_out/s3l_c3_main.json 2> gpurun_out/s3l_c3_main.err
timeout 100 python bench.py --config 4 > gpurun_out/s3l_c4_side.json 2> gpurun_out/s3l_c4_side.err
tail -4 gpurun_out/s3l_tests.log
for f in s3l_c3_side s3l_c3_main s3l_c4_side; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['stages']['launch'][:60])
except Exception as e:
    print('$f', 'failed', e)
PY
done
