#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().splitlines()[-1])
    print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', d['parity']['ok'])
    print(json.dumps(d['other_configs'].get('assembly_others_384_f64')))
    print(json.dumps(d['other_configs'].get('call_latency_us_64cubed_curl')))
    print('asm', d.get('assembly',{}).get('nnz_per_s'), d.get('assembly_c128',{}).get('nnz_per_s'))
except Exception as ex:
    print('parse failed', ex)
PY
