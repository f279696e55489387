#!/bin/bash
mkdir -p gpurun_out /tmp/rep
for c in 4 8 32 64; do
  python bench.py --steps 10 --warmup 3 --no-assembly --no-extras --no-cpu-baseline --e2e-chunks $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunks $c', d['e2e']['ms_per_step'], d['e2e']['value'])" >> gpurun_out/r1l_e2e_chunks.log
done
cat gpurun_out/r1l_e2e_chunks.log
ncu --set full --clock-control none --import-source on -k regex:k_asm -c 1 -o /tmp/rep/asm -f python tools/time_asm.py 512 > gpurun_out/r1l_ncu_asm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o /tmp/rep/fused -f python tools/sweep_fused_te.py > gpurun_out/r1l_ncu_fused.log 2>&1
for r in asm fused; do
  ncu -i /tmp/rep/$r.ncu-rep --page details > gpurun_out/r1l_${r}_details.txt 2>&1
done
grep -E "k_asm|k_fused|Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occ" gpurun_out/r1l_asm_details.txt gpurun_out/r1l_fused_details.txt | head -20
