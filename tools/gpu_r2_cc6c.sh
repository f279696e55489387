#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/sweep_curlcurl.py 256 v6march pec > gpurun_out/r2_sweep_cc6_march.jsonl 2> gpurun_out/r2_sweep_cc6.err; echo "sweep rc=$?"
cat gpurun_out/r2_sweep_cc6_march.jsonl; tail -3 gpurun_out/r2_sweep_cc6.err
