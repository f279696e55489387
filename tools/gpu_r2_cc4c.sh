#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sweep_curlcurl.py 256 full pec > gpurun_out/r2_sweep_cc4_pec.jsonl 2> gpurun_out/r2_sweep_cc4.err; echo "sweep pec rc=$?"
cat gpurun_out/r2_sweep_cc4_pec.jsonl; tail -3 gpurun_out/r2_sweep_cc4.err
timeout 90 python tools/sweep_curlcurl.py 256 quick bloch > gpurun_out/r2_sweep_cc4_bloch.jsonl 2>> gpurun_out/r2_sweep_cc4.err; echo "sweep bloch rc=$?"
cat gpurun_out/r2_sweep_cc4_bloch.jsonl
timeout 500 python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl" --timeout 100 --timeout-method thread > gpurun_out/r2_pytest_cc4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cc4.log
tail -12 gpurun_out/r2_pytest_cc4.log
timeout 300 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_host.py -m gpu -x -q -k "spmv or pml" --timeout 200 --timeout-method thread > gpurun_out/r2_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_new.log
tail -12 gpurun_out/r2_pytest_new.log
timeout 200 python tools/time_spmv.py > gpurun_out/r2_spmv_256.jsonl 2>&1; cat gpurun_out/r2_spmv_256.jsonl | tail -8
