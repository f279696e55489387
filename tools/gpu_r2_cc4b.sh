#!/bin/bash
mkdir -p gpurun_out /tmp/rep
for cfg in "32 6 0 0x34 4" "32 8 0 0x34 4"; do
  set -- $cfg
  tag="ty$2"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cc4 -s 2 -c 1 -o /tmp/rep/cc4_$tag -f \
      python tools/prof_curlcurl.py $1 $2 $3 $(( $4 )) $5 > gpurun_out/r2_ncu_cc4_$tag.log 2>&1
  ncu -i /tmp/rep/cc4_$tag.ncu-rep --page details > gpurun_out/r2_cc4_${tag}_details.txt 2>&1
  ncu -i /tmp/rep/cc4_$tag.ncu-rep --page source --csv > gpurun_out/r2_cc4_${tag}_source.csv 2>&1
  tail -2 gpurun_out/r2_ncu_cc4_$tag.log
done
timeout 600 python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "curlcurl" --timeout 120 --timeout-method thread > gpurun_out/r2_pytest_cc4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cc4.log
tail -12 gpurun_out/r2_pytest_cc4.log
