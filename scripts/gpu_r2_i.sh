#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"; timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_parity_bench_gpu.py tests/test_roll_gpu.py tests/test_modules_gpu.py -q -m gpu 2>&1 | tail -12
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-raster --no-cpu-baseline --kernel-table > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo rc=$?; grep "^kernel" gpurun_out/bench_i.err | head -8; python -c "
import json; d=json.loads(open('gpurun_out/bench_i.json').read()); print(d['value'], d['ms_per_step'])"
echo "=== b32 table"; timeout 600 python bench.py --steps 50 --warmup 5 --batch 32 --no-variants --no-raster --no-cpu-baseline --kernel-table --gemm-table > gpurun_out/bench_i32.json 2> gpurun_out/bench_i32.err; echo rc=$?; grep -E "^kernel|^gemm" gpurun_out/bench_i32.err | head -70; python -c "
import json; d=json.loads(open('gpurun_out/bench_i32.json').read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
