#!/bin/bash
# ncu --set full capture of the Smith-Waterman kernel (one launch of a timed cohort batch) for profiles/r2_classify_full.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"classify_kernel" --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r2_classify_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-grid-stress --from-bam 0 > gpurun_out/r2_classify_full.log 2>&1
ls -la gpurun_out/r2_classify_full.ncu-rep
