#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"; timeout 1200 python -m pytest tests/test_attention_gpu.py tests/test_engine_gpu.py -q -m gpu 2>&1 | tail -12
echo "=== bench bf16"; timeout 600 python bench.py --precision bf16 --steps 20 --warmup 5 --no-variants --no-raster --no-cpu-baseline > gpurun_out/bench_o_bf16.json 2> gpurun_out/bench_o_bf16.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/bench_o_bf16.json').read()); print('bf16', d['value'], d['ms_per_step'])"
