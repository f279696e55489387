#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
tail -4 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
python bench.py --steps 50 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
tail -c 300 gpurun_out/final_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/final_ncu_bench.log 2>&1
python -c "
import json; d=json.loads(open('gpurun_out/final_bench_n1.json').read().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['assembly']['nnz_per_s']/1e9, d['assembly_c128']['nnz_per_s']/1e9, d['cpu_baseline']['value'], d['clocks'])
r=json.loads(open('gpurun_out/final_bench_ref.json').read().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'])"
