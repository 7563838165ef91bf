#!/bin/bash
# 2-GPU check: bench at N=2 (both arms), peer-exchange test, gloo-free
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>gpurun_out/ref2.err | tail -2
echo "=== ours N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_h2.json 2> gpurun_out/bench_h2.err; echo rc=$?; tail -8 gpurun_out/bench_h2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'run', d['run'])
for k in ('tf32_variant','bf16_variant','fp32_variant'):
    v=d.get(k); print(k, v and (v['value'], v['ms_per_step'], v['dp_exchange']))
print('strong', d['strong_scaling'])
PY
echo "=== peer test"; timeout 600 python -m pytest tests/test_adam_nvlink_gpu.py -q -m gpu 2>&1 | tail -3
