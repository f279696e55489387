#!/bin/bash
mkdir -p gpurun_out /tmp/rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cc6 -s 2 -c 1 -o /tmp/rep/cc6 -f \
    python tools/prof_curlcurl.py 0 5 32 3 6 > gpurun_out/r2_ncu_cc6.log 2>&1
ncu -i /tmp/rep/cc6.ncu-rep --page details > gpurun_out/r2_cc6_details.txt 2>&1
ncu -i /tmp/rep/cc6.ncu-rep --page source --csv > gpurun_out/r2_cc6_source.csv 2>&1
tail -2 gpurun_out/r2_ncu_cc6.log
