#!/bin/bash
# One gpurun call: bench, ncu launch list of the bench command, ncu --set full of the top kernels
# (reports are converted to text on the box; only the summaries travel back).
set -x
mkdir -p gpurun_out /tmp/rep
python bench.py --steps 50 --warmup 5 > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_curl3p -s 2 -c 1 -o /tmp/rep/curl3p -f \
    python tools/prof_curl.py 32 4 12 6 4 0 > gpurun_out/r1c_ncu_curl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_asm -c 2 -o /tmp/rep/asm -f \
    python tools/time_asm.py 512 > gpurun_out/r1c_ncu_asm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_term -s 3 -c 1 -o /tmp/rep/term_x -f \
    python tools/time_2d.py 8192 > gpurun_out/r1c_ncu_term.log 2>&1
for r in curl3p asm term_x; do
  ncu -i /tmp/rep/$r.ncu-rep --page details > gpurun_out/r1c_${r}_details.txt 2>&1
  ncu -i /tmp/rep/$r.ncu-rep --page raw --csv > gpurun_out/r1c_${r}_raw.csv 2>&1
  ncu -i /tmp/rep/$r.ncu-rep --page source --csv > gpurun_out/r1c_${r}_source.csv 2>&1
done
python tools/time_2d.py 8192 > gpurun_out/r1c_time_2d.log 2>&1
python tools/time_asm.py > gpurun_out/r1c_time_asm.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt 2>&1
du -sh gpurun_out
