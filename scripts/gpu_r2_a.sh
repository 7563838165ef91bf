#!/bin/bash
# round-2 GPU session A: x3 GEMM check first (bounded), then the full parity suite, smoke, bench lines and kernel tables
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== x3 gemm tests"; timeout 300 python -m pytest tests/test_parity_bench_gpu.py -q -m gpu -k "gemm_tc" -s 2>&1 | tail -40
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 > gpurun_out/pytest_a.log; tail -60 gpurun_out/pytest_a.log
grep -E "deviation|vs oracle|max err" gpurun_out/pytest_a.log | head -80
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench default"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo rc=$?; tail -3 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
echo "=== bench fp32x3 tables"; timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-variants --no-raster --no-cpu-baseline --kernel-table --gemm-table > gpurun_out/bench_a_x3.json 2> gpurun_out/bench_a_x3.err; echo rc=$?; cat gpurun_out/bench_a_x3.err | tail -60; cat gpurun_out/bench_a_x3.json
echo "=== bench tf32 tables"; timeout 600 python bench.py --steps 10 --warmup 3 --no-variants --no-raster --no-cpu-baseline --kernel-table --gemm-table > gpurun_out/bench_a_tf32.json 2> gpurun_out/bench_a_tf32.err; echo rc=$?; cat gpurun_out/bench_a_tf32.err | tail -60
