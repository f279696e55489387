#!/bin/bash
# round-2 multi-GPU check (N = number of visible GPUs): slab parity tests, per-rank timeline, the bench line
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
nproc > gpurun_out/r2_n${N}_host.txt; free -g | head -2 >> gpurun_out/r2_n${N}_host.txt; nvidia-smi topo -m >> gpurun_out/r2_n${N}_host.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/r2_pytest_slab_n${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_slab_n${N}.log
tail -8 gpurun_out/r2_pytest_slab_n${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_timeline.py --steps 200 > gpurun_out/r2_timeline_n${N}.jsonl 2> gpurun_out/r2_timeline_n${N}.err; echo "timeline rc=$?"
cat gpurun_out/r2_timeline_n${N}.jsonl; tail -5 gpurun_out/r2_timeline_n${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}.json').read().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'per_rank', d['per_rank_ms'], 'graph', d['cuda_graph'], 'e2e', d['e2e']['value'])
    print('halo', d['halo_transport'])
    print('parity', d['parity'])
    print('cfg5', d.get('cfg5'))
    print('asm', d.get('assembly'))
except Exception as ex:
    print('parse failed', ex)
PY
