#!/bin/bash
# bench line only, N = number of visible GPUs, the driver's flags
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'per_rank', d['per_rank_ms'], 'graph', d['cuda_graph'], 'e2e', d['e2e']['value'])
    print('clocks', d['clocks'], 'parity', d['parity']['ok'], d['parity']['max_ulp'], d['parity']['cells'])
    print('cfg5', d['cfg5']['value'], d['cfg5']['per_rank_ms'])
    a=d.get('assembly'); print('asm cols', a['ms'], a['nnz_per_s'])
except Exception as ex:
    print('parse failed', ex)
PY
