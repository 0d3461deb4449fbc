#!/bin/bash
# ncu evidence only (one GPU): launch list of two bench steps + `--set full` capture of the 13 launches of one step
TAG=${1:-r02d}
mkdir -p gpurun_out
K='regex:span_kernel|dim_kernel|warp_cols_kernel|lookback_kernel|chain_scan_kernel|chain_where_kernel|pipe_scan_kernel|pipe_where_kernel|warp_scan_cols_kernel|strip_scan_kernel|dim_scan_kernel|where_kernel|fixup_kernel|copy_strided_kernel|fill_kernel|cast_kernel|rbk_|pair_kernel|diff_kernel|expr_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sync-compare --no-check > gpurun_out/b_ncu_$TAG.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches_$TAG.csv)"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" -s 39 -c 13 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-sync-compare --no-check --no-extras > gpurun_out/b_ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
ls -la gpurun_out | grep $TAG
