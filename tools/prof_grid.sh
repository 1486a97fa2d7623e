ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"grid_|finalize|plan_|kde" --csv --log-file gpurun_out/r2_stress_launches.csv python tools/grid_stress.py --problems 64 --readlen 150 --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_stress_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); 
seq=[(r[ik],float(r[iv].replace(',',''))) for r in rows[1:]]
# last call = last 4+ kernels
print('stress: last 12 launches (ns):')
for k,v in seq[-12:]: print('  %-40s %10.0f'%(k[:40],v))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"grid_|finalize|plan_|kde|tally|read_family|fam_|prefilter|unpack|widen" -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_bench_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); 
seq=[(r[ik],float(r[iv].replace(',',''))) for r in rows[1:]]
print('bench: launches 40..75 (ns):')
for k,v in seq[40:75]: print('  %-40s %10.0f'%(k[:40],v))
PY
