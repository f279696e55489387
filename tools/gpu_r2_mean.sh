#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_mean.py 256 > gpurun_out/r2_time_mean.jsonl 2> gpurun_out/r2_time_mean.err; cat gpurun_out/r2_time_mean.jsonl; tail -3 gpurun_out/r2_time_mean.err
timeout 900 python -m pytest tests/test_gpu_matrixfree.py tests/test_gpu_fullsize.py tests/test_gpu_host.py -m gpu -x -q -k "mean or mhat or fullsize or host or device" > gpurun_out/r2_pytest_mean.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mean.log
tail -8 gpurun_out/r2_pytest_mean.log
