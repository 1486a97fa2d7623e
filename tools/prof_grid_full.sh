# full ncu capture of the grid kernels: stress configuration (last call) and the cohort step
ncu --set full --import-source on --clock-control none -k regex:"grid_" --launch-skip 12 --launch-count 4 -o gpurun_out/r2_grid_stress_full -f python tools/grid_stress.py --problems 64 --readlen 150 --reps 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"grid_|finalize|plan_|kde" --launch-skip 42 --launch-count 7 -o gpurun_out/r2_grid_cohort_full -f python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
