#!/bin/bash
# round-2 profiling session: launch list of the default step (with DRAM counters), --set full captures of the new kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches_r2_tf32x3f.csv python bench.py --steps 2 --warmup 3 --no-graph --no-variants --no-raster --no-cpu-baseline > gpurun_out/ncu_r2.log 2>&1
echo "launch list rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_r2_tf32x3f.csv > gpurun_out/launches_r2_tf32x3f.txt 2>&1; head -40 gpurun_out/launches_r2_tf32x3f.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2x3|gemm_tc2b3|gemm_tc2_kernel|attn_tc_fwd|rows_|roll_features" -c 48 \
  -o gpurun_out/r2_kernels python profiles/run_kernels_r2.py > gpurun_out/ncu_r2_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_r2_full.log
python profiles/ncu_summary.py gpurun_out/r2_kernels.ncu-rep > gpurun_out/r2_prof_kernels.txt 2>&1; cat gpurun_out/r2_prof_kernels.txt | cut -c1-260
ls -la gpurun_out/r2_kernels.ncu-rep
