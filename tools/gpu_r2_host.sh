#!/bin/bash
# round-2: host-array entry points (tests + bench e2e through the C ABI) and a source-level ncu capture of the fused curl-of-curl
mkdir -p gpurun_out /tmp/rep
timeout 600 python -m pytest tests/test_gpu_host.py tests/test_gpu_matrixfree.py -m gpu -x -q > gpurun_out/r2_pytest_host.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_host.log
tail -5 gpurun_out/r2_pytest_host.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cfg5 --no-assembly --no-extras > gpurun_out/r2_bench_n1_host.json 2> gpurun_out/r2_bench_n1_host.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1_host.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1_host.json').read().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e'])
except Exception as ex:
    print('parse failed', ex)
PY
ncu --set full --clock-control none --import-source on -k regex:k_curlcurl3 -s 2 -c 1 -o /tmp/rep/cc -f \
    python tools/prof_curlcurl.py 32 7 32 3 > gpurun_out/r2_ncu_cc.log 2>&1
ncu -i /tmp/rep/cc.ncu-rep --page details > gpurun_out/r2_cc_details.txt 2>&1
ncu -i /tmp/rep/cc.ncu-rep --page source --csv > gpurun_out/r2_cc_source.csv 2>&1
tail -2 gpurun_out/r2_ncu_cc.log
