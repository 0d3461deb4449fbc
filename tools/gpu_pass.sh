#!/bin/bash
# GPU pass: parity tests, smoke, bench N=1, ncu launch list + full capture of the hot kernels
# usage: tools/gpu_pass.sh <tag> [skip_ncu]
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_n1_$TAG.err
AFB200_CHAIN_CTAS_PER_SM=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_n1_${TAG}_cta1.json 2>/dev/null
if [ -z "$2" ]; then
K='regex:span_kernel|dim_kernel|warp_cols_kernel|lookback_kernel|chain_scan_kernel|chain_where_kernel|pipe_scan_kernel|pipe_where_kernel|warp_scan_cols_kernel|strip_scan_kernel|where_kernel|scan_fixup_kernel|copy_strided_kernel|fill_kernel|cast_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" -s 39 -c 13 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details --csv > gpurun_out/prof_details_$TAG.csv 2>/dev/null
# the report itself is too large to travel back (gpurun_out is capped at 64 MiB): keep the CSV pages only
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
fi
ls -la gpurun_out
