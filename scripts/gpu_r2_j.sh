#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12
echo "=== b32"; timeout 600 python bench.py --steps 100 --warmup 5 --batch 32 --no-variants --no-raster --no-cpu-baseline > gpurun_out/bench_j32.json 2> gpurun_out/bench_j32.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_j32.json').read()); print('b32', d['value'], d['ms_per_step'])"
