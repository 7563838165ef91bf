#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== module tests"; timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_roll_gpu.py -q -m gpu 2>&1 | tail -8
echo "=== bench default"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo rc=$?; tail -5 gpurun_out/bench_g.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_g.json').read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('from_midi', d['e2e_from_midi'])
print('b32', d['b32']); print('strong', d['strong_scaling'])
PY
