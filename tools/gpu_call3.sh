#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl or fused" > gpurun_out/r1f_pytest_cc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1f_pytest_cc.log
tail -5 gpurun_out/r1f_pytest_cc.log
python tools/sweep_curlcurl.py 256 > gpurun_out/r1f_sweep_curlcurl.log 2>&1
python tools/time_2d.py 8192 > gpurun_out/r1f_time_2d.log 2>&1
tail -3 gpurun_out/r1f_time_2d.log
