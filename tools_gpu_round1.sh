#!/bin/bash
# first GPU pass: parity tests, smoke, bench N=1, ncu launch list + full capture of the hot kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
K='regex:span_kernel|dim_kernel|warp_cols_kernel|lookback_kernel|warp_scan_cols_kernel|strip_scan_kernel|where_kernel|scan_fixup_kernel|copy_strided_kernel|fill_kernel|cast_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" -s 39 -c 13 -o gpurun_out/prof_r01 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out
