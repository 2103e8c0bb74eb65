#!/bin/bash
# one gpurun call: GPU tests, smoke, the bench lines of every config, micro benchmarks (outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/s3_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/s3_gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/s3_smoke.log
timeout 200 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
timeout 100 python bench.py --config 3 > gpurun_out/s3_bench_c3.json 2> gpurun_out/s3_bench_c3.err
timeout 100 python bench.py --config 4 > gpurun_out/s3_bench_c4.json 2> gpurun_out/s3_bench_c4.err
timeout 100 python bench.py --config 5 > gpurun_out/s3_bench_c5.json 2> gpurun_out/s3_bench_c5.err
timeout 60 python benchmarks/attn_decode.py > gpurun_out/s3_attn.log 2>&1
timeout 60 python benchmarks/dcn_bwd.py > gpurun_out/s3_dcnbwd.log 2>&1
tail -3 gpurun_out/s3_gputests.log; cut -c1-600 gpurun_out/s3_bench.json
