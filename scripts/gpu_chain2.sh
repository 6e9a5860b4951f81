#!/bin/bash
# chain kernel: full parity suite, bench, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -6
for chain in 1 0; do
  echo -n "PIFU_CHAIN=$chain  "
  PIFU_CHAIN=$chain timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_chain$chain.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.4g q/s  ms/step %.2f  e2e %.4g  %s: TF %.1f frac %.3f share %.3f avg_launch_us %.1f launches %d clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel'][:12], r['achieved'], r['frac'], r['share_of_step'], r['avg_launch_us'], d['gpu_launches'], d['clocks']))"
done
# launch list (cold-cache, serialised: shares only): two column chunks of the 256^3 lattice
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chain.csv python scripts/profile_step.py 2424832 > gpurun_out/prof_launch.log 2>&1
tail -1 gpurun_out/prof_launch.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_chain.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+1:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value')
tot={}
for r in data:
    k=r[ki].split('(')[0].split('::')[-1][:40]; tot[k]=tot.get(k,0)+float(r[vi])/1000
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:8]: print('%-42s %10.1f us' % (k,v))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/chain_r01 -f python scripts/profile_step.py 2424832 > gpurun_out/prof_full.log 2>&1
tail -2 gpurun_out/prof_full.log
ls -la gpurun_out | head -20
