#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/sweep_curlcurl.py 256 v6 pec > gpurun_out/r2_sweep_cc6_pec.jsonl 2> gpurun_out/r2_sweep_cc6.err; echo "sweep pec rc=$?"
cat gpurun_out/r2_sweep_cc6_pec.jsonl; tail -3 gpurun_out/r2_sweep_cc6.err
timeout 500 python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl" --timeout 100 --timeout-method thread > gpurun_out/r2_pytest_cc6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cc6.log
tail -6 gpurun_out/r2_pytest_cc6.log
