#!/bin/bash
mkdir -p gpurun_out /tmp/rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cc6 -s 2 -c 1 -o /tmp/rep/cc6 -f \
    python tools/prof_curlcurl.py 0 0 0 0 0 > gpurun_out/r2_ncu_cc6.log 2>&1
ncu -i /tmp/rep/cc6.ncu-rep --page details > gpurun_out/r2_cc6_details.txt 2>&1
ncu -i /tmp/rep/cc6.ncu-rep --page source --csv > gpurun_out/r2_cc6_source.csv 2>&1
ncu -i /tmp/rep/cc6.ncu-rep --page raw --csv > gpurun_out/r2_cc6_raw.csv 2>&1
tail -2 gpurun_out/r2_ncu_cc6.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_asm_warp -s 1 -c 2 -o /tmp/rep/asm -f python tools/time_asm.py 512 > gpurun_out/r2_ncu_asm.log 2>&1
ncu -i /tmp/rep/asm.ncu-rep --page details > gpurun_out/r2_asm_details.txt 2>&1
ncu -i /tmp/rep/asm.ncu-rep --page raw --csv > gpurun_out/r2_asm_raw.csv 2>&1
tail -2 gpurun_out/r2_ncu_asm.log
