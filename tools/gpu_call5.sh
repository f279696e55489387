#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j_pytest_gpu.log
tail -4 gpurun_out/r1j_pytest_gpu.log
python tools/time_2d.py 8192 > gpurun_out/r1j_time_2d.log 2>&1
cat gpurun_out/r1j_time_2d.log
