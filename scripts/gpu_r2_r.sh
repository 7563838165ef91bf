#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== attention tests"; timeout 600 python -m pytest tests/test_attention_gpu.py -q -m gpu 2>&1 | tail -3
echo "=== parity"; timeout 900 python -m pytest tests/test_parity_bench_gpu.py -q -m gpu -k "forward_vs_oracle or dropout or replay" 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-raster --no-cpu-baseline --kernel-table > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo rc=$?; grep "^kernel" gpurun_out/bench_r.err | head -6; python -c "
import json; d=json.loads(open('gpurun_out/bench_r.json').read()); print(d['value'], d['ms_per_step'])"
