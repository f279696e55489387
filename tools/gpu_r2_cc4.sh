#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl" > gpurun_out/r2_pytest_cc4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cc4.log
tail -15 gpurun_out/r2_pytest_cc4.log
timeout 300 python tools/sweep_curlcurl.py 256 ${1:-full} > gpurun_out/r2_sweep_cc4.jsonl 2> gpurun_out/r2_sweep_cc4.err; echo "sweep rc=$?"
cat gpurun_out/r2_sweep_cc4.jsonl; tail -3 gpurun_out/r2_sweep_cc4.err
