#!/bin/bash
# Standard on-box sequence (run through gpurun): GPU parity tests, one bench line, an ncu launch list.
#   scripts/gpu_check.sh <tag> [quick]
# Outputs land in gpurun_out/ (merged back into the repo by gpurun).
tag=${1:-run}
mode=${2:-full}
mkdir -p gpurun_out
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "pytest rc=$?"
if [ "$mode" = "quick" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-raster > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
else
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
fi
echo "bench rc=$?"
python - <<EOF
import json
d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", {k: r[k] for k in r if k not in ("kernel", "peak_source", "tensor")}, "tensor TF/s", r["tensor"]["achieved"])
print("raster", (d.get("rasteriser") or {}).get("roofline"))
print("cpu", d.get("cpu_baseline"))
EOF
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-raster > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_$tag.csv | head -40
