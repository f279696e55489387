#!/bin/bash
# round-2 single-GPU check: GPU tests, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt; free -g | head -2 >> gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().splitlines()[-1])
    print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', d['parity'])
    print('cfg5', d.get('cfg5'))
    print('asm', d.get('assembly',{}).get('nnz_per_s'), d.get('assembly_c128',{}).get('nnz_per_s'))
    print('other', json.dumps(d.get('other_configs'))[:1500])
except Exception as ex:
    print('parse failed', ex)
PY
