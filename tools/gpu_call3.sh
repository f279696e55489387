#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl" > gpurun_out/r1k_pytest_cc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1k_pytest_cc.log
tail -5 gpurun_out/r1k_pytest_cc.log
python tools/sweep_curlcurl.py 256 > gpurun_out/r1k_sweep_curlcurl.log 2>&1
