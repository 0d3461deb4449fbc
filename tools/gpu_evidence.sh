#!/bin/bash
# evidence pass (one GPU): bench N=1, ncu launch list + full capture of the hot kernels, compute-sanitizer memcheck / racecheck
# usage: tools/gpu_evidence.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1_$TAG.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sync", d["sync_api"]["value"], "parity", d["parity_checked"], "cpu", d["cpu_baseline"])
for k,v in d["ops"].items(): print("   %-20s %8.4f ms %8.1f GB/s %.3f" % (k, v["ms"], v["GBps"], v["frac_of_peak_per_gpu"]))
for k,v in (d.get("extras") or {}).items(): print("   %-50s %s" % (k, v))
PY
K='regex:span_kernel|dim_kernel|warp_cols_kernel|lookback_kernel|chain_scan_kernel|chain_where_kernel|pipe_scan_kernel|pipe_where_kernel|warp_scan_cols_kernel|strip_scan_kernel|dim_scan_kernel|where_kernel|fixup_kernel|copy_strided_kernel|fill_kernel|cast_kernel|rbk_|pair_kernel|diff_kernel|expr_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sync-compare --no-check > gpurun_out/b_ncu_$TAG.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches_$TAG.csv)"
# full capture: the 13 launches of the last timed step (3 warm-up + 1 step = 52 launches before the extras)
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" -s 39 -c 13 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-sync-compare --no-check --no-extras > gpurun_out/b_ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
# compute-sanitizer over the golden-vector tests (small shapes, every kernel family) -- memcheck, then racecheck
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py tests/test_gpu_async.py -m gpu -x -q --timeout 1200 > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck_$TAG.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py -m gpu -x -q --timeout 1200 > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck_$TAG.log
tail -5 gpurun_out/sanitizer_memcheck_$TAG.log; tail -5 gpurun_out/sanitizer_racecheck_$TAG.log
ls -la gpurun_out | tail -12
