#!/bin/bash
# the driver's own sequence: GPU tests, smoke, default bench, reference arm, ncu launch list of the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().splitlines()[-1])
    print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity']['ok'])
    print('asm', d.get('assembly',{}).get('nnz_per_s'), d.get('assembly_c128',{}).get('nnz_per_s'), 'cfg5', d.get('cfg5',{}).get('value'))
    print('line bytes', len(open('gpurun_out/r2_bench_n1.json').read()))
    r=json.loads(open('gpurun_out/r2_bench_ref.json').read().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'])
except Exception as ex:
    print('parse failed', ex)
PY
