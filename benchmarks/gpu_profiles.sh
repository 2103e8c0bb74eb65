#!/bin/bash
# profiles: launch lists (cold-cache, serialised) + ncu --set full of the new kernels
mkdir -p gpurun_out
export MR_BENCH_SKIP_CPU=1
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 5000 --csv --log-file gpurun_out/r2f_cfg3_launches.csv python bench.py --config 3 --steps 1 --warmup 1 > gpurun_out/r2f_cfg3_ncu.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum -c 9000 --csv --log-file gpurun_out/r2f_cfg5_launches.csv python bench.py --config 5 --steps 1 --warmup 1 > gpurun_out/r2f_cfg5_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/attn_fwd_r2 python benchmarks/attn_train_once.py > gpurun_out/r2f_attn_fwd.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:attn_bwd -s 1 -c 1 -o gpurun_out/attn_bwd_r2 python benchmarks/attn_train_once.py > gpurun_out/r2f_attn_bwd.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:dcn_dgrad -s 1 -c 1 -o gpurun_out/dcn_dgrad_r2 python benchmarks/dcn_bwd_once.py > gpurun_out/r2f_dcn_dgrad.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:dcn_wgrad -s 1 -c 1 -o gpurun_out/dcn_wgrad_r2 python benchmarks/dcn_bwd_once.py > gpurun_out/r2f_dcn_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2f_*.csv
