#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest_gpu.log
python tools/time_2d.py 8192 > gpurun_out/r1e_time_2d.log 2>&1
python tools/time_asm.py > gpurun_out/r1e_time_asm.log 2>&1
tail -5 gpurun_out/r1e_pytest_gpu.log
