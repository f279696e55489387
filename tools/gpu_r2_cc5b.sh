#!/bin/bash
mkdir -p gpurun_out /tmp/rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cc5 -s 2 -c 1 -o /tmp/rep/cc5 -f \
    python tools/prof_curlcurl.py 0 5 32 4 5 > gpurun_out/r2_ncu_cc5.log 2>&1
ncu -i /tmp/rep/cc5.ncu-rep --page details > gpurun_out/r2_cc5_details.txt 2>&1
ncu -i /tmp/rep/cc5.ncu-rep --page source --csv > gpurun_out/r2_cc5_source.csv 2>&1
tail -2 gpurun_out/r2_ncu_cc5.log
