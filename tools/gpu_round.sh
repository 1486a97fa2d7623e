#!/bin/bash
# One GPU-box pass: parity tests, per-run counters behind the rooflines (ncu), the bench line, the launch list.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_$tag.log 2>&1; tail -3 gpurun_out/r2_gputest_$tag.log
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fp64.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:"classify_kernel|grid_" --csv --log-file gpurun_out/r2_cal.csv \
    python tools/calibrate.py run gpurun_out/r2_cal_stats.json > gpurun_out/r2_cal.log 2>&1
tail -2 gpurun_out/r2_cal.log
python tools/calibrate.py merge gpurun_out/r2_cal.csv gpurun_out/r2_cal_stats.json > /dev/null && cp profiles/r2_counters.json gpurun_out/r2_counters.json
python bench.py > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; tail -c 600 gpurun_out/r2_bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_$tag.json 2>> gpurun_out/r2_bench_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --from-bam 0 > /dev/null 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_$tag.json"))
print(json.dumps({k:d[k] for k in ("value","ms_per_step","e2e","roofline","roofline_grid","roofline_grid_stress","from_bam") if k in d})[:3000])
PY
