#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sweep_curlcurl.py 256 full pec > gpurun_out/r2_sweep_cc5_pec.jsonl 2> gpurun_out/r2_sweep_cc5.err; echo "sweep pec rc=$?"
cat gpurun_out/r2_sweep_cc5_pec.jsonl; tail -3 gpurun_out/r2_sweep_cc5.err
timeout 500 python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl" --timeout 100 --timeout-method thread > gpurun_out/r2_pytest_cc5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cc5.log
tail -6 gpurun_out/r2_pytest_cc5.log
