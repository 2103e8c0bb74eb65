#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_decode_gpu.py tests/test_attention_train_gpu.py tests/test_conv_engine_gpu.py tests/test_trunks_engine_gpu.py tests/test_surfaces_gpu.py -q > gpurun_out/s3j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3j_tests.log
tail -4 gpurun_out/s3j_tests.log
