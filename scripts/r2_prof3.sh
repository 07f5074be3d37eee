#!/bin/bash
mkdir -p gpurun_out
python -W ignore scripts/profile_report.py cfg5 0.1 > gpurun_out/r2_prof3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"reassign_kernel|cc_count_kernel|cc_hist_kernel|cc_compact_kernel|select_fill_kernel|select_count_kernel|tile_build_kernel|emit_tiles_kernel" -c 24 -o gpurun_out/r2c_report_kernels -f python -W ignore scripts/profile_report.py cfg5 0.1 > gpurun_out/r2_prof3_ncu.log 2>&1
cat gpurun_out/r2_prof3_plain.log
for cfg in "cfg4 0.05" "cfg3 0.2"; do
  echo "== $cfg"
  timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "layout|fit:|parity|OK|rror|assert|Trace"
done
