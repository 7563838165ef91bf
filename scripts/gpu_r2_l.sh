#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== b3 gemm tests"; timeout 300 python -m pytest tests/test_parity_bench_gpu.py -q -m gpu -k "b3" -s 2>&1 | grep -E "bf16x3 M|passed|failed|FAILED|Error|assert " | head -30
echo "=== bench bf16x3f"; timeout 600 python bench.py --precision bf16x3f --steps 20 --warmup 5 --no-variants --no-raster --no-cpu-baseline --gemm-table > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; echo rc=$?; grep -E "^gemm bf16x3" gpurun_out/bench_l.err | head -5; python -c "
import json; d=json.loads(open('gpurun_out/bench_l.json').read()); print(d['value'], d['ms_per_step'])"
