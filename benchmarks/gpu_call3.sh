#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_attention_decode_gpu.py tests/test_attention_train_gpu.py -q > gpurun_out/s3c_attn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3c_attn_tests.log
timeout 500 python -m pytest tests -m gpu -q --deselect tests/test_attention_decode_gpu.py --deselect tests/test_attention_train_gpu.py > gpurun_out/s3c_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/s3c_gputests.log
for c in 3 4 5; do
  timeout 150 python bench.py --config $c > gpurun_out/s3c_bench_c$c.json 2> gpurun_out/s3c_bench_c$c.err
done
timeout 80 python benchmarks/attn_decode.py > gpurun_out/s3c_attn_decode.log 2>&1
timeout 120 python benchmarks/attn_train.py > gpurun_out/s3c_attn_train.log 2>&1
tail -5 gpurun_out/s3c_attn_tests.log; tail -5 gpurun_out/s3c_gputests.log
cat gpurun_out/s3c_attn_decode.log gpurun_out/s3c_attn_train.log | tail -12
for c in 3 4 5; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s3c_bench_c$c.json').read().strip().splitlines()[-1])
    print($c, d['ms_per_step'], d['value'], d['e2e']['value'], d['stages']['launch'])
except Exception as e:
    print($c, 'failed', e)
PY
done
