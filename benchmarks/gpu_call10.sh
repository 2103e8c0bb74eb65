#!/bin/bash
mkdir -p gpurun_out
export MR_BENCH_SKIP_CPU=1
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum -c 4200 --csv --log-file gpurun_out/r2g_cfg3_launches.csv python bench.py --config 3 --steps 1 --warmup 1 > gpurun_out/r2g_cfg3_ncu.log 2>&1
ls -la gpurun_out/r2g_cfg3_launches.csv
