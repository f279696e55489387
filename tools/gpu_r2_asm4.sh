#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_asm_others.py 384 > gpurun_out/r2_time_asm_cells.jsonl 2> gpurun_out/r2_time_asm_cells.err; cat gpurun_out/r2_time_asm_cells.jsonl; tail -3 gpurun_out/r2_time_asm_cells.err
timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not spmv" > gpurun_out/r2_pytest_asm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_asm.log
tail -5 gpurun_out/r2_pytest_asm.log
