#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_decode_gpu.py tests/test_attention_train_gpu.py tests/test_surfaces_gpu.py -q > gpurun_out/s3d_attn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3d_attn_tests.log
timeout 80 python benchmarks/attn_decode.py > gpurun_out/s3d_attn_decode.log 2>&1
timeout 120 python benchmarks/attn_train.py > gpurun_out/s3d_attn_train.log 2>&1
timeout 150 python bench.py --config 4 > gpurun_out/s3d_bench_c4.json 2> gpurun_out/s3d_bench_c4.err
tail -5 gpurun_out/s3d_attn_tests.log
cat gpurun_out/s3d_attn_decode.log gpurun_out/s3d_attn_train.log | tail -12
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s3d_bench_c4.json').read().strip().splitlines()[-1])
    print(4, d['ms_per_step'], d['value'], d['e2e']['value'], d['stages']['launch'])
except Exception as e:
    print(4, 'failed', e)
PY
