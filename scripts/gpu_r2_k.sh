#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== b3 gemm tests"; timeout 300 python -m pytest tests/test_parity_bench_gpu.py -q -m gpu -k "b3" -s 2>&1 | grep -E "bf16x3 M|passed|failed|FAILED|Error|assert " | head -30
echo "=== parity bf16x3f"; timeout 900 python -m pytest tests/test_parity_bench_gpu.py tests/test_sanitize_gpu.py -q -m gpu -k "bf16x3f" 2>&1 | tail -8
echo "=== diag"; timeout 900 python profiles/micro/diag_precision.py 2>&1 | grep -E "^tf32 |^tf32x3f  |^bf16x3f" 
echo "=== bench bf16x3f"; timeout 600 python bench.py --precision bf16x3f --steps 20 --warmup 5 --no-variants --no-raster --no-cpu-baseline --kernel-table --gemm-table > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo rc=$?; grep -E "^kernel|^gemm bf16x3" gpurun_out/bench_k.err | head -24; python -c "
import json; d=json.loads(open('gpurun_out/bench_k.json').read()); print(d['value'], d['ms_per_step'])"
