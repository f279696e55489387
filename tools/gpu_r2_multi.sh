#!/bin/bash
START=$(date +%s)
# multi-GPU check (N = number of visible GPUs): slab parity tests, the bench line, host<->device bandwidth with all ranks copying
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
[ "$SKIP_SLAB" = 1 ] || timeout 300 python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/r2_pytest_slab_n${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_slab_n${N}.log
tail -4 gpurun_out/r2_pytest_slab_n${N}.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n${N}.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/pcie_bw.py > gpurun_out/r2_pcie_n${N}.json 2> gpurun_out/r2_pcie_n${N}.err; echo "pcie rc=$?"
cat gpurun_out/r2_pcie_n${N}.json; tail -2 gpurun_out/r2_pcie_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'per_rank', d['per_rank_ms'], 'graph', d['cuda_graph'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('parity', d['parity']['ok'], d['parity']['max_ulp'], d['parity']['cells'])
    print('cfg5', d['cfg5']['value'], d['cfg5']['per_rank_ms'])
    a=d.get('assembly'); print('asm cols', a['ms'], a['nnz_per_s'], 'rows', a['row_partition']['ms'], a['row_partition']['nnz_per_s'])
except Exception as ex:
    print('parse failed', ex)
PY
echo "elapsed $(( $(date +%s) - START )) s"
