#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/r1i_bench_n1.json 2> gpurun_out/r1i_bench_n1.err
tail -c 600 gpurun_out/r1i_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r1i_bench_n1.json'))
print(d['value'], d['e2e'], d['assembly']['ms'], d['assembly_c128'].get('ms'), json.dumps(d['other_configs'])[:1500])"
