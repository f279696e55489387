#!/bin/bash
mkdir -p gpurun_out /tmp/rep
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
timeout 300 python tools/time_asm.py 512 > gpurun_out/r2_time_asm.jsonl 2>&1; cat gpurun_out/r2_time_asm.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_asm_group -s 1 -c 2 -o /tmp/rep/asm -f python tools/time_asm.py 512 > gpurun_out/r2_ncu_asm.log 2>&1
ncu -i /tmp/rep/asm.ncu-rep --page details > gpurun_out/r2_asm_details.txt 2>&1
ncu -i /tmp/rep/asm.ncu-rep --page source --csv > gpurun_out/r2_asm_source.csv 2>&1
tail -2 gpurun_out/r2_ncu_asm.log
