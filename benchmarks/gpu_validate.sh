#!/bin/bash
# one gpurun call that validates a revision: all GPU tests, smoke(), the bench line of every configuration (outputs under gpurun_out/)
#   /usr/local/graft/bin/gpurun --timeout 1800 -- "bash benchmarks/gpu_validate.sh"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/s3f_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/s3f_gputests.log
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/s3f_smoke.log
timeout 300 python bench.py > gpurun_out/s3f_bench.json 2> gpurun_out/s3f_bench.err
for c in 3 4 5; do
  timeout 150 python bench.py --config $c > gpurun_out/s3f_bench_c$c.json 2> gpurun_out/s3f_bench_c$c.err
done
tail -8 gpurun_out/s3f_gputests.log; tail -3 gpurun_out/s3f_smoke.log
cut -c1-400 gpurun_out/s3f_bench.json
for c in 3 4 5; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s3f_bench_c$c.json').read().strip().splitlines()[-1])
    print($c, d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d['stages']['launch'])
except Exception as e:
    print($c, 'failed', e)
PY
done
