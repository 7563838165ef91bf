#!/bin/bash
# N-GPU bench line (both arms are the driver's job; here: ours) -> gpurun_out/bench_r2_${N}gpu.json
N=$1
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo rc=$?; tail -3 gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_${N}gpu.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'identical', d['run']['ranks_identical'], d['run']['dp_exchange'][:40])
for k in ('tf32_variant','bf16_variant','fp32_variant'):
    v=d.get(k); print(k, v and (v['value'], v['ms_per_step']))
print('strong', [(s['global_batch'], s['value'], s['ms_per_step']) for s in d['strong_scaling'] or []])
print('clocks', d['clocks'])
PY
