#!/bin/bash
# evidence run after the long-row attention work: full GPU suite, smoke, default bench line (+ tables), the config 3 sweep,
# ncu of the long-row kernels (T = 129 and 257)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench default"; timeout 900 python bench.py --kernel-table --gemm-table > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo rc=$?; grep -E "^kernel" gpurun_out/bench_t.err | head -14
echo "=== sweep"; timeout 900 python bench.py --mode sweep > gpurun_out/sweep_t.jsonl 2> gpurun_out/sweep_t.err; echo rc=$?
for T in 129 257; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcl --launch-skip 4 --launch-count 4 -o gpurun_out/attn_long_r2c_$T -f \
    python profiles/micro/prof_attn_long.py $T > gpurun_out/ncu_long_$T.log 2>&1; echo "ncu $T rc=$?"
done
