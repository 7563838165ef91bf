#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2x3|gemm_tc2b3" -s 2 -c 2 -o /tmp/x3src python profiles/run_gemm_x3_only.py > gpurun_out/ncu_src.log 2>&1
echo "rc=$?"
ncu -i /tmp/x3src.ncu-rep --page source --csv > gpurun_out/x3_source.csv 2> gpurun_out/x3_source.err
ls -la gpurun_out/x3_source.csv; head -c 600 gpurun_out/x3_source.csv
ncu -i /tmp/x3src.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=[h for h in hdr if any(k in h for k in ("smsp__average_warp","smsp__pcsamp_warps_issue_stalled","l1tex__data_pipe_lsu_wavefronts_mem_shared","sm__inst_executed_pipe_tensor","smsp__warp_issue_stalled","sm__pipe_tensor","lsu_mem_shared"))]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:40])
    for h in want[:60]:
        print("   ", h, r[hdr.index(h)])
PY
