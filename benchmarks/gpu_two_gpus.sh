#!/bin/bash
# 2-GPU checks of the data-parallel paths (one process per GPU, NCCL): headline bench and the graph-captured cfg-3 step
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MR_BENCH_SKIP_CPU=1 timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s3h_bench_n2.json 2> gpurun_out/s3h_bench_n2.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --config 3 > gpurun_out/s3h_bench_c3_n2.json 2> gpurun_out/s3h_bench_c3_n2.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --config 4 > gpurun_out/s3h_bench_c4_n2.json 2> gpurun_out/s3h_bench_c4_n2.err
for f in s3h_bench_n2 s3h_bench_c3_n2 s3h_bench_c4_n2; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('final_loss'), d.get('stages',{}).get('launch'))
except Exception as e:
    print('$f', 'failed', e)
PY
done
tail -5 gpurun_out/s3h_bench_c3_n2.err
