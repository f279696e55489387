#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matrixfree.py -m gpu -x -q -k "fused or divg or curl" > gpurun_out/r1o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1o_pytest.log
tail -6 gpurun_out/r1o_pytest.log
python bench.py --steps 20 --warmup 3 --no-assembly --no-cpu-baseline > gpurun_out/r1o_bench.json 2>gpurun_out/r1o_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1o_bench.json').read().splitlines()[-1]); o=d['other_configs']; print(d['value']); print(o.get('error')); print(o['divg_3d_c128'], o['grad_3d_c128'], o['curlcurl_fused'])"
