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
# ncu launch list of the timed step (same command as the round-1 list): the step is two launches of k_curl3p
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-assembly --no-extras --no-cfg5 > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2_bench_launches.csv') if not l.startswith('==')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: v = float(r[ix['Metric Value']].replace(',', ''))
    except Exception: continue
    k = r[ix['Kernel Name']][:70]; tot[k] += v; cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(6): print(f"{100*v/s:5.1f}%  n={cnt[k]:4d}  avg={v/cnt[k]/1e3:8.1f} us  {k}")
PY
