#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== gemm tests"; timeout 300 python -m pytest tests/test_parity_bench_gpu.py -q -m gpu -k "gemm" 2>&1 | tail -3
for prec in tf32x3f bf16x3f fp32x3; do
echo "=== bench $prec"; timeout 600 python bench.py --precision $prec --steps 20 --warmup 5 --no-variants --no-raster --no-cpu-baseline --gemm-table > gpurun_out/bench_m_$prec.json 2> gpurun_out/bench_m_$prec.err; echo rc=$?; grep -E "^gemm (bf16x3|tf32x3) M=133120" gpurun_out/bench_m_$prec.err | head -6; python -c "
import json; d=json.loads(open('gpurun_out/bench_m_$prec.json').read()); print(d['value'], d['ms_per_step'])"
done
