#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/r2_pytest_asm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_asm.log
tail -5 gpurun_out/r2_pytest_asm.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().splitlines()[-1])
    print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e'])
    print('asm', d.get('assembly',{}).get('nnz_per_s'), d.get('assembly_c128',{}).get('nnz_per_s'))
except Exception as ex:
    print('parse failed', ex)
PY
