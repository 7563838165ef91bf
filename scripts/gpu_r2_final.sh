#!/bin/bash
# final evidence run of a build: full GPU suite, smoke, default bench line (+ tables), reference arm
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> /dev/null; cat gpurun_out/bench_final_reference.json | cut -c1-200
echo "=== bench default"; timeout 900 python bench.py --kernel-table --gemm-table > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?; grep -E "^kernel" gpurun_out/bench_final.err | head -14
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
r=d['roofline']; print('roofline', r['kernel'][:30], r['bound'], r['achieved'], r['frac'], r.get('executed_frac'))
print('other', [(o['kernel'][:24], o['bound'], round(o['achieved']), round(o['frac'],3)) for o in d['roofline_other']])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k in ('tf32_variant','bf16_variant','fp32_variant','b32'):
    v=d.get(k); print(k, v and (round(v['value']), round(v['ms_per_step'],3)))
print('strong', [(s['global_batch'], round(s['value']), round(s['ms_per_step'],3)) for s in d['strong_scaling'] or []])
print('midi', d['e2e_from_midi']['value'], d['e2e_from_midi']['stage_ms'])
print('raster', d['rasteriser']['roofline']['frac'], d['rasteriser']['ms'])
print('clocks', d['clocks'])
PY
