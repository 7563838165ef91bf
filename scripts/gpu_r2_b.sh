#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== targeted tests"; timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_bench_gpu.py -q -m gpu 2>&1 | tail -25
echo "=== precision diag"; timeout 900 python profiles/micro/diag_precision.py 2>&1 | tail -12
echo "=== bench tf32 tables"; timeout 600 python bench.py --steps 10 --warmup 3 --no-variants --no-raster --no-cpu-baseline --kernel-table > gpurun_out/bench_b_tf32.json 2> gpurun_out/bench_b_tf32.err; echo rc=$?; grep "^kernel" gpurun_out/bench_b_tf32.err | head -12; python -c "
import json; d=json.loads(open('gpurun_out/bench_b_tf32.json').read()); print(d['value'], d['ms_per_step'])"
