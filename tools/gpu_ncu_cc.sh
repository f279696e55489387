#!/bin/bash
mkdir -p gpurun_out /tmp/rep
ncu --set full --clock-control none --import-source on -k regex:k_curlcurl3 -s 2 -c 1 -o /tmp/rep/cc -f \
    python tools/prof_curlcurl.py $1 $2 $3 $4 > gpurun_out/r1g_ncu_cc.log 2>&1
ncu -i /tmp/rep/cc.ncu-rep --page details > gpurun_out/r1g_cc_details.txt 2>&1
ncu -i /tmp/rep/cc.ncu-rep --page raw --csv > gpurun_out/r1g_cc_raw.csv 2>&1
ncu -i /tmp/rep/cc.ncu-rep --page source --csv > gpurun_out/r1g_cc_source.csv 2>&1
tail -2 gpurun_out/r1g_ncu_cc.log
