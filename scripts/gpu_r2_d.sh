#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench default"; timeout 900 python bench.py --steps 20 --warmup 5 --kernel-table --gemm-table > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo rc=$?; grep -E "^kernel|^gemm" gpurun_out/bench_d.err | head -50; cat gpurun_out/bench_d.json
